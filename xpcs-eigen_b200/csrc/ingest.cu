// ingest.cu -- Filter stage on the device.
//
// Replaces, for the whole frame range at once, the per-frame work of
//   Imm::NextFrames            (reference io/imm.cpp:70-118)
//   SparseFilter::Apply        (reference filter/sparse_filter.cpp:115-193)
//   DenseFilter::Apply         (reference filter/dense_filter.cpp:121-210)
//   DarkImage::Compute         (reference data_structure/dark_image.cpp:81-106)
// and builds the pixel-major store that plays the role of data_structure::SparseData
// (reference data_structure/sparse_data.cpp:59-103).
//
// Store layout: unmasked pixels ("rows") are numbered in (dq, sq, pixel) order; 32
// consecutive rows form a slice; a slice of length L is L*32 words, word (j, lane) at
// slice_base + j*32 + lane, so that one warp owns one slice, every lane streams its own
// row with fully coalesced 128-byte accesses and shared-memory staging is conflict free.
// Word formats: kPacked  uint32  (frame << 12) | count     (exact integer path)
//               kFloat   uint64  (frame << 32) | float bits
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "internal.h"

namespace xpcs {

// ------------------------------------------------------------------------------------
// frame lookup helpers (sparse IMM source: events are concatenated per raw frame)
// ------------------------------------------------------------------------------------

// first raw frame of every block of kEvPerBlock events: largest f with off[f] <= e
__global__ void k_block_frames(const int64_t *__restrict__ off, int nraw, int64_t E,
                               int *__restrict__ first, int nblocks, int64_t ev_begin)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    int64_t e = (int64_t)b * kEvPerBlock + (ev_begin & ~3LL);
    int lo = 0, hi = nraw;  // invariant: off[lo] <= e, answer in [lo, hi)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= e) lo = mid;
        else hi = mid;
    }
    first[b] = lo;
}

struct IngestArgs {
    const int32_t *idx;
    const int16_t *val;
    const int64_t *off;
    int nraw;
    int64_t E;
    const int *block_first;
    const int *row_of_pixel;
    const double *flat;
    int *row_count;
    double *frame_acc;
    long long *summary;
    const int64_t *slice_base;
    void *store;
    int F, stride, rawblock, P;
    // dense source (explicit frame id + float value per event)
    const int32_t *evt;
    const float *valf;
    // stream scatter: slice-ordered records
    unsigned long long *slice_end;  // per slice: end of its record stream, counted down (one atomic gives the position)
    unsigned long long *rec;
    // chunked ingest: the events [ev_begin, E) of idx/val, off points at the chunk's first frame
    int64_t ev_begin;  // first event of the chunk (the blocks start at ev_begin & ~3: aligned vector loads)
    int frame_base;    // raw frame number of off[0]
};

// summary slots
enum { kSumBadCount = 0, kSumMaxLen = 1, kSumEvents = 2, kSumOverflow = 3, kSumWords = 4, kSumMaxCount = 5, kSumSlots = 8 };

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Maps a raw frame to its output frame (-1 = not used): blocks of rawblock = stride*avg raw
// frames, every stride-th one taken (reference main.cpp:258-268, sparse_filter.cpp:143).
__device__ __forceinline__ int out_frame(int raw, int rawblock, int stride, int F)
{
    if (rawblock == 1) return raw < F ? raw : -1;
    int t = raw / rawblock;
    int sub = raw - t * rawblock;
    if (t >= F || (sub % stride) != 0) return -1;
    return t;
}

// Pass 1 over the frame-major events: per-row histogram, per-frame sums, range check.
template <int KIND, bool DENSE_SRC>
__global__ void __launch_bounds__(kIngestThreads) k_hist(IngestArgs a)
{
    const int64_t base = (a.ev_begin & ~3LL) + (int64_t)blockIdx.x * kEvPerBlock;
    int f = DENSE_SRC ? 0 : a.block_first[blockIdx.x];
    constexpr int kIter = kEvPerBlock / (kIngestThreads * 4);
#pragma unroll 1
    for (int k = 0; k < kIter; k++) {
        const int64_t e0 = base + ((int64_t)k * kIngestThreads + threadIdx.x) * 4;
        int pix[4], t[4];
        double v[4];
        int nv = 0;
        if (e0 < a.E) nv = (a.E - e0) >= 4 ? 4 : (int)(a.E - e0);
        int16_t raw[4] = {0, 0, 0, 0};
        float rawf[4] = {0.f, 0.f, 0.f, 0.f};
        if (nv == 4) {
            int4 p4 = *reinterpret_cast<const int4 *>(a.idx + e0);
            pix[0] = p4.x; pix[1] = p4.y; pix[2] = p4.z; pix[3] = p4.w;
            if (DENSE_SRC) {
                int4 t4 = *reinterpret_cast<const int4 *>(a.evt + e0);
                t[0] = t4.x; t[1] = t4.y; t[2] = t4.z; t[3] = t4.w;
                float4 f4 = *reinterpret_cast<const float4 *>(a.valf + e0);
                rawf[0] = f4.x; rawf[1] = f4.y; rawf[2] = f4.z; rawf[3] = f4.w;
            } else {
                short4 s4 = *reinterpret_cast<const short4 *>(a.val + e0);
                raw[0] = s4.x; raw[1] = s4.y; raw[2] = s4.z; raw[3] = s4.w;
            }
        } else {
            for (int j = 0; j < nv; j++) {
                pix[j] = a.idx[e0 + j];
                if (DENSE_SRC) { t[j] = a.evt[e0 + j]; rawf[j] = a.valf[e0 + j]; }
                else raw[j] = a.val[e0 + j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            v[j] = 0.0;
            if (j >= nv || e0 + j < a.ev_begin) { t[j] = -1; continue; }
            if (!DENSE_SRC) {
                const int64_t e = e0 + j;
                while (f + 1 < a.nraw && e >= __ldg(a.off + f + 1)) f++;
                t[j] = out_frame(f + a.frame_base, a.rawblock, a.stride, a.F);
            }
            int r = -1;
            if (t[j] >= 0 && (unsigned)pix[j] < (unsigned)a.P) r = __ldg(a.row_of_pixel + pix[j]);
            if (r < 0) { t[j] = -1; continue; }
            atomicAdd(a.row_count + r, 1);
            if (DENSE_SRC) v[j] = (double)rawf[j];
            else if (KIND == kPacked) {
                v[j] = (double)raw[j];
                if (raw[j] < 0 || raw[j] >= (1 << kCountBits)) a.summary[kSumBadCount] = 1;
            } else {
                // reference sparse_filter.cpp:152: float v = value[j] * flatfield_[pix]
                v[j] = (double)(float)((double)(float)raw[j] * __ldg(a.flat + pix[j]));
            }
        }
        if (DENSE_SRC) continue;  // the dense filter already accumulated the frame sums
        // per-output-frame sums; one atomic per warp when the whole warp sits in one frame
        int tl = -1;
        bool same = true;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (t[j] < 0) continue;
            if (tl < 0) tl = t[j];
            else if (t[j] != tl) same = false;
            s += v[j];
        }
        int tref = __reduce_max_sync(0xffffffffu, tl);
        bool uniform = __all_sync(0xffffffffu, same && (tl < 0 || tl == tref));
        if (uniform) {
            double w = warp_sum(s);
            if ((threadIdx.x & 31) == 0 && tref >= 0) atomicAdd(a.frame_acc + tref, w);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (t[j] >= 0) atomicAdd(a.frame_acc + t[j], v[j]);
        }
    }
}

// slice length = longest row of the slice; also copies the histogram to row_len.
__global__ void k_slice_len(const int *__restrict__ row_count, int *__restrict__ row_len,
                            int *__restrict__ slice_len, int *__restrict__ slice_cur, int n_slices,
                            long long *summary)
{
    int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (s >= n_slices) return;
    int c = row_count[s * kSlice + lane];
    row_len[s * kSlice + lane] = c;
    int m = __reduce_max_sync(0xffffffffu, c);
    int tot = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) {
        slice_len[s] = m;
        slice_cur[s] = tot;
        atomicMax(summary + kSumMaxLen, (long long)m);
        atomicAdd((unsigned long long *)summary + kSumEvents, (unsigned long long)tot);
    }
}

// exclusive scan of slice_len*32 over all slices (single CTA; a few 10^4..10^5 entries)
// blockIdx.x == 1: the same scan over the slices' event totals -> first record of every slice
__global__ void __launch_bounds__(1024) k_slice_scan(const int *__restrict__ slice_len,
                                                     int64_t *__restrict__ slice_base,
                                                     const int *__restrict__ slice_tot,
                                                     int64_t *__restrict__ slice_rec,
                                                     unsigned long long *__restrict__ slice_end,
                                                     int n_slices, long long *summary)
{
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const int per = (n_slices + 1023) / 1024;
    const int a = tid * per, b = min(n_slices, a + per);
    if (blockIdx.x == 1) {
        long long s = 0;
        for (int i = a; i < b; i++) s += (long long)slice_tot[i];
        part[tid] = s;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            long long x = tid >= o ? part[tid - o] : 0;
            __syncthreads();
            part[tid] += x;
            __syncthreads();
        }
        long long run = part[tid] - s;
        for (int i = a; i < b; i++) {
            slice_rec[i] = run;
            run += (long long)slice_tot[i];
            slice_end[i] = (unsigned long long)run;
        }
        if (tid == 1023) slice_rec[n_slices] = part[1023];
        return;
    }
    long long s = 0;
    for (int i = a; i < b; i++) s += (long long)slice_len[i] * kSlice;
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        long long x = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += x;
        __syncthreads();
    }
    long long run = part[tid] - s;
    for (int i = a; i < b; i++) {
        slice_base[i] = run;
        run += (long long)slice_len[i] * kSlice;
    }
    if (tid == 1023) {
        slice_base[n_slices] = part[1023];
        summary[kSumWords] = part[1023];
    }
}

// Pass 2 over the frame-major events: scatter into the slices.  The rank inside a row is
// taken from the histogram counting down, so the store needs no zeroed cursor and the
// histogram is all zeros again afterwards (ready for the next ingest).
template <int KIND, bool DENSE_SRC>
__global__ void __launch_bounds__(kIngestThreads) k_scatter(IngestArgs a)
{
    const int64_t base = (a.ev_begin & ~3LL) + (int64_t)blockIdx.x * kEvPerBlock;
    int f = DENSE_SRC ? 0 : a.block_first[blockIdx.x];
    constexpr int kIter = kEvPerBlock / (kIngestThreads * 4);
#pragma unroll 1
    for (int k = 0; k < kIter; k++) {
        const int64_t e0 = base + ((int64_t)k * kIngestThreads + threadIdx.x) * 4;
        int nv = 0;
        if (e0 < a.E) nv = (a.E - e0) >= 4 ? 4 : (int)(a.E - e0);
        int pix[4], t[4];
        int16_t raw[4] = {0, 0, 0, 0};
        float rawf[4] = {0.f, 0.f, 0.f, 0.f};
        if (nv == 4) {
            int4 p4 = *reinterpret_cast<const int4 *>(a.idx + e0);
            pix[0] = p4.x; pix[1] = p4.y; pix[2] = p4.z; pix[3] = p4.w;
            if (DENSE_SRC) {
                int4 t4 = *reinterpret_cast<const int4 *>(a.evt + e0);
                t[0] = t4.x; t[1] = t4.y; t[2] = t4.z; t[3] = t4.w;
                float4 f4 = *reinterpret_cast<const float4 *>(a.valf + e0);
                rawf[0] = f4.x; rawf[1] = f4.y; rawf[2] = f4.z; rawf[3] = f4.w;
            } else {
                short4 s4 = *reinterpret_cast<const short4 *>(a.val + e0);
                raw[0] = s4.x; raw[1] = s4.y; raw[2] = s4.z; raw[3] = s4.w;
            }
        } else {
            for (int j = 0; j < nv; j++) {
                pix[j] = a.idx[e0 + j];
                if (DENSE_SRC) { t[j] = a.evt[e0 + j]; rawf[j] = a.valf[e0 + j]; }
                else raw[j] = a.val[e0 + j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j >= nv || e0 + j < a.ev_begin) continue;
            int tj;
            if (DENSE_SRC) tj = t[j];
            else {
                const int64_t e = e0 + j;
                while (f + 1 < a.nraw && e >= __ldg(a.off + f + 1)) f++;
                tj = out_frame(f + a.frame_base, a.rawblock, a.stride, a.F);
            }
            if (tj < 0 || (unsigned)pix[j] >= (unsigned)a.P) continue;
            const int r = __ldg(a.row_of_pixel + pix[j]);
            if (r < 0) continue;
            const int rank = atomicSub(a.row_count + r, 1) - 1;
            const int64_t dst = __ldg(a.slice_base + (r >> 5)) + (int64_t)rank * kSlice + (r & 31);
            if (KIND == kPacked) {
                reinterpret_cast<uint32_t *>(a.store)[dst] =
                    ((uint32_t)tj << kCountBits) | (uint32_t)(raw[j] & ((1 << kCountBits) - 1));
            } else {
                float v = DENSE_SRC ? rawf[j]
                                    : (float)((double)(float)raw[j] * __ldg(a.flat + pix[j]));
                reinterpret_cast<unsigned long long *>(a.store)[dst] =
                    ((unsigned long long)(uint32_t)tj << 32) | (unsigned long long)__float_as_uint(v);
            }
        }
    }
}

template <int KIND>
struct WordT;
template <>
struct WordT<kPacked> {
    typedef uint32_t type;
};
template <>
struct WordT<kFloat> {
    typedef unsigned long long type;
};

// Stream scatter (the default): the direct scatter above writes 4- or 8-byte words 128 bytes apart,
// and a 128-byte line of the store is complete only when all 32 rows of the slice have received that
// event rank -- with ~10^5 slices the open lines outgrow the L2 and DRAM sees partial-sector
// traffic (3.5x the algorithmic bytes measured).  Here an event is appended to the record stream
// of its SLICE instead (one open line per slice), and k_place -- one CTA per slice, the slice's
// region of the store hot in L2 -- moves the records to their rows.
// Record: kPacked  lane << 32 | word;  kFloat  lane << 59 | frame << 32 | float bits (frames < 2^27).
constexpr int kRecLaneShiftF = 59;

template <int KIND, bool DENSE_SRC>
__global__ void __launch_bounds__(kIngestThreads) k_scatter_rec(IngestArgs a)
{
    const int64_t base = (a.ev_begin & ~3LL) + (int64_t)blockIdx.x * kEvPerBlock;
    int f = DENSE_SRC ? 0 : a.block_first[blockIdx.x];
    constexpr int kIter = kEvPerBlock / (kIngestThreads * 4);
#pragma unroll 1
    for (int k = 0; k < kIter; k++) {
        const int64_t e0 = base + ((int64_t)k * kIngestThreads + threadIdx.x) * 4;
        int nv = 0;
        if (e0 < a.E) nv = (a.E - e0) >= 4 ? 4 : (int)(a.E - e0);
        int pix[4], t[4];
        int16_t raw[4] = {0, 0, 0, 0};
        float rawf[4] = {0.f, 0.f, 0.f, 0.f};
        if (nv == 4) {
            int4 p4 = *reinterpret_cast<const int4 *>(a.idx + e0);
            pix[0] = p4.x; pix[1] = p4.y; pix[2] = p4.z; pix[3] = p4.w;
            if (DENSE_SRC) {
                int4 t4 = *reinterpret_cast<const int4 *>(a.evt + e0);
                t[0] = t4.x; t[1] = t4.y; t[2] = t4.z; t[3] = t4.w;
                float4 f4 = *reinterpret_cast<const float4 *>(a.valf + e0);
                rawf[0] = f4.x; rawf[1] = f4.y; rawf[2] = f4.z; rawf[3] = f4.w;
            } else {
                short4 s4 = *reinterpret_cast<const short4 *>(a.val + e0);
                raw[0] = s4.x; raw[1] = s4.y; raw[2] = s4.z; raw[3] = s4.w;
            }
        } else {
            for (int j = 0; j < nv; j++) {
                pix[j] = a.idx[e0 + j];
                if (DENSE_SRC) { t[j] = a.evt[e0 + j]; rawf[j] = a.valf[e0 + j]; }
                else raw[j] = a.val[e0 + j];
            }
        }
        // three rounds over the thread's four events, so that the four row look-ups and then the four returning
        // atomics are in flight together (the kernel is bound by the latency of the atomics, not by their count)
        int rr[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            rr[j] = -1;
            if (j >= nv || e0 + j < a.ev_begin) continue;
            if (!DENSE_SRC) {
                const int64_t e = e0 + j;
                while (f + 1 < a.nraw && e >= __ldg(a.off + f + 1)) f++;
                t[j] = out_frame(f + a.frame_base, a.rawblock, a.stride, a.F);
            }
            if (t[j] < 0 || (unsigned)pix[j] >= (unsigned)a.P) continue;
            rr[j] = __ldg(a.row_of_pixel + pix[j]);
        }
        unsigned long long pos[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (rr[j] >= 0) pos[j] = atomicAdd(a.slice_end + (rr[j] >> 5), ~0ull) - 1ull;  // -1: the stream fills from its end
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (rr[j] < 0) continue;
            const int r = rr[j], tj = t[j];
            unsigned long long rec;
            if (KIND == kPacked) {
                rec = ((unsigned long long)(r & 31) << 32) |
                      (unsigned long long)(((uint32_t)tj << kCountBits) | (uint32_t)(raw[j] & ((1 << kCountBits) - 1)));
            } else {
                float v = DENSE_SRC ? rawf[j]
                                    : (float)((double)(float)raw[j] * __ldg(a.flat + pix[j]));
                rec = ((unsigned long long)(r & 31) << kRecLaneShiftF) | ((unsigned long long)(uint32_t)tj << 32) |
                      (unsigned long long)__float_as_uint(v);
            }
            a.rec[pos[j]] = rec;
        }
    }
}

// One CTA per slice: records -> rows, in stream order (a stable placement: the rank of a record is the
// number of earlier records of the same row -- 32 running cursors, plus the records of the row in the
// warps before this one in the 512-record step, plus a match inside the warp).  The stream was filled
// from its end, so reading it forwards fills the rows latest-first and nearly sorted, which is what the
// finalisation kernels expect from the scatter.  A whole CTA per slice keeps the number of slices in
// flight (and with it the partially written lines of the store) within the L2.
constexpr int kPlaceWarps = 16;

template <int KIND>
__global__ void __launch_bounds__(kPlaceWarps * 32) k_place(const unsigned long long *__restrict__ rec,
                                                            const int64_t *__restrict__ slice_rec,
                                                            const int64_t *__restrict__ slice_base, void *store)
{
    typedef typename WordT<KIND>::type W;
    __shared__ int cur[kSlice];
    __shared__ int wcnt[kPlaceWarps][kSlice];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x;
    if (warp == 0) cur[lane] = 0;
    const int64_t r0 = slice_rec[s];
    const int n = (int)(slice_rec[s + 1] - r0);
    W *dst = reinterpret_cast<W *>(store) + slice_base[s];
    for (int e0 = 0; e0 < n; e0 += kPlaceWarps * 32) {
        const int e = e0 + (int)threadIdx.x;
        const bool ok = e < n;
        const unsigned long long x = ok ? rec[r0 + e] : 0ull;
        int row = 32 + lane;  // idle lanes match nobody
        W w = 0;
        if (ok) {
            if (KIND == kPacked) {
                row = (int)(x >> 32);
                w = (W)(uint32_t)x;
            } else {
                row = (int)(x >> kRecLaneShiftF);
                w = (W)(x & ((1ull << kRecLaneShiftF) - 1ull));
            }
        }
        wcnt[warp][lane] = 0;
        __syncwarp();
        const unsigned peers = __match_any_sync(0xffffffffu, row);
        const int before = __popc(peers & ((1u << lane) - 1u));
        if (ok && before == 0) wcnt[warp][row] = __popc(peers);
        __syncthreads();
        if (warp == 0) {  // exclusive prefix over the warps, per row; the cursors move on
            int run = cur[lane];
#pragma unroll
            for (int k = 0; k < kPlaceWarps; k++) {
                const int c = wcnt[k][lane];
                wcnt[k][lane] = run;
                run += c;
            }
            cur[lane] = run;
        }
        __syncthreads();
        if (ok) dst[(int64_t)(wcnt[warp][row] + before) * kSlice + row] = w;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------
// Row finalisation: one warp per slice, one lane per row.  Sorts the row by frame, merges
// events that fell on the same output frame (duplicates, stride/average blocks), applies
// /avg and the frame-sum normalisation, and produces pixelSum and the per-static-bin sums
// of sparse_filter.cpp:176-185 from the finished rows.
// ------------------------------------------------------------------------------------
struct FinalizeArgs {
    void *store;
    const int64_t *slice_base;
    const int *slice_len;
    int *row_len;
    const int *sbin_of_row;
    double *row_sum;
    double *part_total;
    double *part_partial;
    const float *frame_scale;  // nullptr unless normalize_by_framesum
    long long *summary;
    int n_slices, S, swindow, avg, smem_len;
    unsigned char *flagged;  // warp-per-row kernel: slices it leaves; lane-per-row kernel: only those (nullptr = all)
    int row_cap, window;     // warp-per-row kernel: longest row it takes, first ranking window
    // lane-per-row kernel, every slice in shared memory: the rows are taken straight from the slice's
    // record stream (stream scatter), k_place is skipped
    const unsigned long long *rec;
    const int64_t *slice_rec;
    int accumulate;          // chunked ingest: row_sum += instead of =
    int late_window;         // XPCS_COMPAT_LATE_WINDOW: frame t > 0 belongs to static window (t - 1) / swindow
};

template <int KIND>
__global__ void __launch_bounds__(32) k_finalize(FinalizeArgs a)
{
    typedef typename WordT<KIND>::type W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int curS[kSlice];
    const int s = blockIdx.x;
    const int lane = threadIdx.x;
    const int r = s * kSlice + lane;
    if (a.flagged && !a.flagged[s]) return;  // done by k_finalize_warp
    const int len = a.slice_len[s];
    if (len == 0) {
        if (!a.accumulate) a.row_sum[r] = 0.0;
        return;
    }
    W *g = reinterpret_cast<W *>(a.store) + a.slice_base[s] + lane;
    const int n = a.row_len[r];
    const bool in_smem = len <= a.smem_len;
    W *col;
    if (in_smem && a.rec) {
        // stable placement of the slice's records (see k_place); the stream runs latest-first, so a
        // record of rank k goes to position n - 1 - k of its row and the column is close to ascending
        W *tile = reinterpret_cast<W *>(smem_raw);
        col = tile + lane;
        const int64_t r0 = a.slice_rec[s];
        const int nrec = (int)(a.slice_rec[s + 1] - r0);
        curS[lane] = 0;  // records of every row placed so far
        __syncwarp();
        unsigned long long x1 = lane < nrec ? a.rec[r0 + lane] : 0ull;
        unsigned long long x2 = 32 + lane < nrec ? a.rec[r0 + 32 + lane] : 0ull;
        for (int e0 = 0; e0 < nrec; e0 += 32) {
            const int e = e0 + lane;
            const bool ok = e < nrec;
            const unsigned long long x = x1;
            x1 = x2;
            x2 = e + 64 < nrec ? a.rec[r0 + e + 64] : 0ull;
            int row = 32 + lane;  // idle lanes match nobody
            W w = 0;
            if (ok) {
                if (KIND == kPacked) {
                    row = (int)(x >> 32);
                    w = (W)(uint32_t)x;
                } else {
                    row = (int)(x >> kRecLaneShiftF);
                    w = (W)(x & ((1ull << kRecLaneShiftF) - 1ull));
                }
            }
            const unsigned peers = __match_any_sync(0xffffffffu, row);
            const int before = __popc(peers & ((1u << lane) - 1u));
            int rank = 0;
            if (ok) rank = curS[row] + before;
            __syncwarp();
            if (ok && before == 0) curS[row] += __popc(peers);
            __syncwarp();
            const int nrow = __shfl_sync(0xffffffffu, n, row & 31);
            if (ok) tile[(nrow - 1 - rank) * kSlice + row] = w;
        }
        __syncwarp();
    } else if (in_smem) {
        col = reinterpret_cast<W *>(smem_raw) + lane;
        // the scatter filled the row from the top down: reverse while loading so that the
        // column is close to ascending already
        for (int j = 0; j < len; j++)
            if (j < n) col[(n - 1 - j) * kSlice] = g[(int64_t)j * kSlice];
    } else {
        col = g;
        for (int i = 0, j = n - 1; i < j; i++, j--) {
            W x = col[(int64_t)i * kSlice];
            col[(int64_t)i * kSlice] = col[(int64_t)j * kSlice];
            col[(int64_t)j * kSlice] = x;
        }
    }
    // insertion sort by frame (the key sits in the high bits of the word); the largest word so far stays in a
    // register, so that a row that is sorted already costs one load and one compare per event
    W top = n > 0 ? col[0] : (W)0;
    for (int i = 1; i < n; i++) {
        W w = col[(int64_t)i * kSlice];
        if (w >= top) {
            top = w;
            continue;
        }
        int j = i;
        while (j > 0) {
            W p = col[(int64_t)(j - 1) * kSlice];
            if (p <= w) break;
            col[(int64_t)j * kSlice] = p;
            j--;
        }
        col[(int64_t)j * kSlice] = w;
    }
    // merge equal frames
    int m = 0;
    if (KIND == kPacked) {
        uint32_t prev = 0xffffffffu;
        for (int i = 0; i < n; i++) {
            uint32_t w = (uint32_t)col[(int64_t)i * kSlice];
            uint32_t key = w >> kCountBits;
            if (m > 0 && key == prev) {
                uint32_t q = (uint32_t)col[(int64_t)(m - 1) * kSlice];
                uint32_t c = (q & ((1u << kCountBits) - 1)) + (w & ((1u << kCountBits) - 1));
                if (c >= (1u << kCountBits)) a.summary[kSumOverflow] = 1;
                col[(int64_t)(m - 1) * kSlice] = (W)((key << kCountBits) | (c & ((1u << kCountBits) - 1)));
            } else {
                col[(int64_t)m * kSlice] = (W)w;
                m++;
            }
            prev = key;
        }
    } else {
        uint32_t prev = 0xffffffffu;
        for (int i = 0; i < n; i++) {
            unsigned long long w = (unsigned long long)col[(int64_t)i * kSlice];
            uint32_t key = (uint32_t)(w >> 32);
            if (m > 0 && key == prev) {
                unsigned long long q = (unsigned long long)col[(int64_t)(m - 1) * kSlice];
                float c = __fadd_rn(__uint_as_float((uint32_t)q), __uint_as_float((uint32_t)w));
                col[(int64_t)(m - 1) * kSlice] = (W)(((unsigned long long)key << 32) | __float_as_uint(c));
            } else {
                col[(int64_t)m * kSlice] = (W)w;
                m++;
            }
            prev = key;
        }
    }
    // sums in frame order (pixels_sum_ += v is sequential fp32 in the reference)
    const int sb = a.sbin_of_row[r];
    double total;
    int cmax = 0;
    {
        float fsum = 0.0f;
        long long isum = 0;
        // current static window and its first frame beyond; window w holds the frames [w * swindow + late, (w + 1) *
        // swindow + late) (frame 0 belongs to window 0 either way)
        int win = 0, wend = a.swindow + (a.late_window ? 1 : 0);
        double wacc = 0.0;
        bool open = false;
        for (int i = 0; i < m; i++) {
            int t;
            double v;
            if (KIND == kPacked) {
                uint32_t w = (uint32_t)col[(int64_t)i * kSlice];
                t = (int)(w >> kCountBits);
                int c = (int)(w & ((1u << kCountBits) - 1));
                isum += c;
                cmax = max(cmax, c);
                v = (double)c;
            } else {
                unsigned long long w = (unsigned long long)col[(int64_t)i * kSlice];
                t = (int)(w >> 32);
                float x = __uint_as_float((uint32_t)w);
                if (a.avg > 1) x = __fdiv_rn(x, (float)a.avg);
                fsum = __fadd_rn(fsum, x);
                v = (double)x;
                if (a.frame_scale) x = __fdiv_rn(x, a.frame_scale[t]);
                col[(int64_t)i * kSlice] = (W)(((unsigned long long)(uint32_t)t << 32) | __float_as_uint(x));
            }
            if (t >= wend) {
                // frames ascend.  The lanes of the warp cross their window borders at different events, so whatever
                // stands here runs in nearly every iteration: step to the next window, no integer division (which was
                // 19 % of this kernel's instructions)
                if (open && sb >= 0) atomicAdd(a.part_partial + (int64_t)win * a.S + sb, wacc);
                do {
                    win++;
                    wend += a.swindow;
                } while (t >= wend);
                wacc = 0.0;
                open = false;
            }
            wacc += v;
            open = true;
        }
        if (open && sb >= 0) atomicAdd(a.part_partial + (int64_t)win * a.S + sb, wacc);
        total = (KIND == kPacked) ? (double)isum : (double)fsum;
    }
    if (a.accumulate) a.row_sum[r] += total;
    else a.row_sum[r] = total;
    if (sb >= 0 && m > 0) atomicAdd(a.part_total + sb, total);
    a.row_len[r] = m;
    if (KIND == kPacked && cmax > 2048) atomicMax(a.summary + kSumMaxCount, (long long)cmax);  // rare: fp16 two-time operand
    if (in_smem) {
        __syncwarp();
        for (int j = 0; j < len; j++)
            if (j < m) g[(int64_t)j * kSlice] = col[j * kSlice];
    }
}

// ---- warp-per-row finalisation (long rows) ---------------------------------------------------
// Same job as k_finalize, organised for rows of hundreds of events, where a whole slice per
// lane-per-row warp no longer fits shared memory at a useful occupancy: a CTA owns a slice,
// each warp takes rows in turn into a private double buffer.  The scatter leaves a row nearly
// sorted (atomics of concurrently running CTAs only swap neighbours), so the sort is a ranking
// pass over a +-window neighbourhood, position = index - (larger words before) + (smaller words
// after), checked afterwards (the target buffer is pre-filled with a sentinel: a collision
// leaves a hole, a hole or an inversion fails the check) and redone over the whole row when the
// window was too small -- with the whole row as window the formula is the stable rank, always a
// permutation.  Merging, /avg, frame-sum scaling and the sums run lanes-over-events with
// ballot compaction and a segmented scan per static window.
constexpr int kFwWarps = 16;

template <int KIND>
__global__ void __launch_bounds__(kFwWarps * 32) k_finalize_warp(FinalizeArgs a)
{
    typedef typename WordT<KIND>::type W;
    extern __shared__ __align__(16) unsigned char fw_smem[];
    const int s = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int len = a.slice_len[s];
    if (len > a.row_cap) {  // CTA-uniform: the lane-per-row kernel works this slice in global memory
        if (tid == 0) a.flagged[s] = 1;
        return;
    }
    if (len == 0) {
        if (tid < kSlice && !a.accumulate) a.row_sum[s * kSlice + tid] = 0.0;
        return;
    }
    const W kMaxW = ~(W)0;
    W *bufA = reinterpret_cast<W *>(fw_smem) + (size_t)warp * 2 * a.row_cap;
    W *bufB = bufA + a.row_cap;
    constexpr int kShift = KIND == kPacked ? kCountBits : 32;

    for (int rr = warp; rr < kSlice; rr += kFwWarps) {
        const int r = s * kSlice + rr;
        const int n = a.row_len[r];
        W *g = reinterpret_cast<W *>(a.store) + a.slice_base[s] + rr;
        W *A = bufA, *B = bufB;
        // the scatter filled the row from the top down: reverse while loading so that it is close
        // to ascending already
        for (int j = lane; j < n; j += 32) A[n - 1 - j] = g[(int64_t)j * kSlice];
        __syncwarp();
        // ---- sort by (frame, value bits)
        bool sorted = true;
        for (int i = lane; i + 1 < n; i += 32) sorted = sorted && !(A[i] > A[i + 1]);
        if (!__all_sync(0xffffffffu, sorted)) {
            int D = n > 4 * a.window ? a.window : n;
            for (;;) {
                for (int j = lane; j < n; j += 32) B[j] = kMaxW;
                __syncwarp();
                for (int c0 = 0; c0 < n; c0 += 32) {
                    const int i = c0 + lane;
                    if (i < n) {
                        const W key = A[i];
                        int pos = i;
                        const int lo = max(0, i - D), hi = min(n - 1, i + D);
                        for (int j = lo; j < i; j++) pos -= A[j] > key ? 1 : 0;
                        for (int j = i + 1; j <= hi; j++) pos += A[j] < key ? 1 : 0;
                        B[pos] = key;
                    }
                }
                __syncwarp();
                bool ok = true;
                for (int i = lane; i < n; i += 32) ok = ok && B[i] != kMaxW && (i + 1 >= n || !(B[i] > B[i + 1]));
                if (__all_sync(0xffffffffu, ok) || D >= n) break;
                D = 4 * D < n ? 4 * D : n;  // (the displacement grows with the GPU count: fewer rows share the same scatter CTAs)
                __syncwarp();
            }
            W *t = A;
            A = B;
            B = t;
        }
        __syncwarp();
        // ---- merge equal frames, /avg, frame-sum scaling, sums; compacted row goes to B
        const int sb = a.sbin_of_row[r];
        int m_base = 0;
        double tot = 0.0;
        long long itot = 0;
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int i = c0 + lane;
            const bool valid = i < n;
            const W w = valid ? A[i] : kMaxW;
            const uint32_t key = (uint32_t)(w >> kShift);
            const bool head = valid && (i == 0 || (uint32_t)(A[i - 1] >> kShift) != key);
            W outw = w;
            double v = 0.0;
            if (head) {
                if (KIND == kPacked) {
                    uint32_t c = (uint32_t)w & ((1u << kCountBits) - 1u);
                    for (int k = i + 1; k < n && (uint32_t)(A[k] >> kShift) == key; k++)
                        c += (uint32_t)A[k] & ((1u << kCountBits) - 1u);
                    if (c >= (1u << kCountBits)) a.summary[kSumOverflow] = 1;
                    outw = (W)((key << kCountBits) | (c & ((1u << kCountBits) - 1u)));
                    itot += c;
                    v = (double)c;
                    if (c > 2048u) atomicMax(a.summary + kSumMaxCount, (long long)c);  // rare: fp16 two-time operand
                } else {
                    float x = __uint_as_float((uint32_t)w);
                    for (int k = i + 1; k < n && (uint32_t)(A[k] >> kShift) == key; k++)
                        x = __fadd_rn(x, __uint_as_float((uint32_t)A[k]));
                    if (a.avg > 1) x = __fdiv_rn(x, (float)a.avg);
                    tot += (double)x;
                    v = (double)x;
                    if (a.frame_scale) x = __fdiv_rn(x, a.frame_scale[key]);
                    outw = (W)(((unsigned long long)key << 32) | (unsigned long long)__float_as_uint(x));
                }
            }
            const unsigned mk = __ballot_sync(0xffffffffu, head);
            if (head) B[m_base + __popc(mk & ((1u << lane) - 1u))] = outw;
            m_base += __popc(mk);
            // per-static-window sums: the frames ascend, so a window is a run of lanes
            int wi = 0x7fffffff;
            if (valid) wi = a.late_window ? (key > 0u ? (int)((key - 1u) / (uint32_t)a.swindow) : 0) : (int)(key / (uint32_t)a.swindow);
            double x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, x, o);
                const int wo = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o && wo == wi) x += y;
            }
            const int wn = __shfl_down_sync(0xffffffffu, wi, 1);
            if (valid && sb >= 0 && (lane == 31 || wn != wi) && x != 0.0)
                atomicAdd(a.part_partial + (int64_t)wi * a.S + sb, x);
        }
        __syncwarp();
        for (int j = lane; j < m_base; j += 32) g[(int64_t)j * kSlice] = B[j];
        double total;
        if (KIND == kPacked) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) itot += __shfl_xor_sync(0xffffffffu, itot, o);
            total = (double)itot;
        } else {
            total = warp_sum(tot);
            total = (double)(float)total;
        }
        if (lane == 0) {
            if (a.accumulate) a.row_sum[r] += total;
            else a.row_sum[r] = total;
            a.row_len[r] = m_base;
            if (sb >= 0 && m_base > 0) atomicAdd(a.part_total + sb, total);
        }
        __syncwarp();
    }
}

// frame_scale[t] = frameSum[t] / mean(frameSum)   (reference main.cpp:313-337)
__global__ void k_frame_scale(const double *__restrict__ frame_acc, float *__restrict__ scale,
                              int F, int P)
{
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < F; i += blockDim.x)
        s += (double)__fdiv_rn((float)frame_acc[i], (float)P);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const float mean = __fdiv_rn((float)red[0], (float)F);
    for (int i = threadIdx.x; i < F; i += blockDim.x)
        scale[i] = __fdiv_rn(__fdiv_rn((float)frame_acc[i], (float)P), mean);
}

// ------------------------------------------------------------------------------------
// Dark image and dense filter
// ------------------------------------------------------------------------------------

// One thread per pixel, frames in order: the running mean / M2 recurrences of
// dark_image.cpp:88-101 in fp64, without FMA contraction, then sqrt(M2/n).
__global__ void k_dark(const int16_t *__restrict__ frames, int n, int P,
                       const double *__restrict__ flat, double *__restrict__ avg,
                       double *__restrict__ sd)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P) return;
    double a = 0.0, m2 = 0.0;
    const double fl = flat[j];
    for (int i = 0; i < n; i++) {
        double before = a;
        double pix = __dmul_rn((double)(float)frames[(int64_t)i * P + j], fl);
        a = __dadd_rn(a, __ddiv_rn(__dsub_rn(pix, a), (double)(i + 1)));
        m2 = __dadd_rn(m2, __dmul_rn(__dsub_rn(pix, before), __dsub_rn(pix, a)));
    }
    avg[j] = a;
    sd[j] = sqrt(__ddiv_rn(m2, (double)n));
}

struct DenseArgs {
    const int16_t *frames;  // [nframes][P]
    int first_raw, nframes, P, F, stride, rawblock;
    const int *row_of_pixel;
    const double *flat, *dark_avg, *dark_std;  // dark_* nullptr without darks
    float lld, sigma;
    int32_t *out_idx;
    int32_t *out_t;
    float *out_v;
    unsigned long long *counter;
    unsigned long long capacity;
    double *frame_acc;
    long long *summary;
};

// dense_filter.cpp:150-174 for a batch of raw frames.  One CTA owns 512 pixels (two per thread)
// and walks kDfFrames consecutive frames: the per-pixel constants (mask, dark average,
// threshold, flat-field) are loaded once into registers instead of once per frame, the frames
// are read as coalesced 4-byte pairs, survivors are staged in shared memory and appended to the
// frame-major event list (pixel, output frame, value) with ONE global atomic per flush and
// coalesced stores, and the per-frame sums leave the CTA as one atomic per frame.
constexpr int kDfThreads = 256;
constexpr int kDfPixels = 2 * kDfThreads;
constexpr int kDfFrames = 64;
constexpr int kDfCap = 3072;        // staged events per CTA
constexpr int kDfFlushEvery = 4;    // frames between fill checks

__global__ void __launch_bounds__(kDfThreads) k_dense_filter(DenseArgs a)
{
    __shared__ int s_pix[kDfCap];
    __shared__ int s_t[kDfCap];
    __shared__ float s_v[kDfCap];
    __shared__ double s_fsum[kDfFrames];
    __shared__ int s_count;
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x;
    const int j0 = (blockIdx.x * kDfThreads + tid) * 2;
    const int f_begin = blockIdx.y * kDfFrames;
    const int f_end = min(a.nframes, f_begin + kDfFrames);
    if (tid == 0) s_count = 0;
    for (int i = tid; i < kDfFrames; i += kDfThreads) s_fsum[i] = 0.0;
    // per-pixel constants
    bool valid[2];
    double davg[2] = {0.0, 0.0}, flat[2] = {1.0, 1.0};
    float thresh[2] = {0.0f, 0.0f};
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int j = j0 + k;
        valid[k] = j < a.P && a.row_of_pixel[j] >= 0;
        if (valid[k]) {
            flat[k] = a.flat[j];
            if (a.dark_avg) {
                davg[k] = a.dark_avg[j];
                thresh[k] = (float)__dadd_rn((double)a.lld, __dmul_rn((double)a.sigma, a.dark_std[j]));
            }
        }
    }
    const bool pair_ok = (a.P & 1) == 0 && j0 + 1 < a.P;  // aligned 4-byte loads
    __syncthreads();

    auto flush = [&]() {  // all threads
        __syncthreads();
        const int n = min(s_count, kDfCap);
        if (tid == 0) s_base = atomicAdd(a.counter, (unsigned long long)n);
        __syncthreads();
        const unsigned long long base = s_base;
        for (int i = tid; i < n; i += kDfThreads) {
            const unsigned long long pos = base + i;
            if (pos < a.capacity) {
                a.out_idx[pos] = s_pix[i];
                a.out_t[pos] = s_t[i];
                a.out_v[pos] = s_v[i];
            } else a.summary[kSumOverflow] = 2;
        }
        __syncthreads();
        if (tid == 0) s_count = 0;
        __syncthreads();
    };

    for (int fr = f_begin; fr < f_end; fr++) {
        const int t = out_frame(a.first_raw + fr, a.rawblock, a.stride, a.F);
        if (t >= 0) {
            const int16_t *src = a.frames + (int64_t)fr * a.P;
            short raw[2] = {0, 0};
            if (pair_ok) {
                const short2 r2 = *reinterpret_cast<const short2 *>(src + j0);
                raw[0] = r2.x;
                raw[1] = r2.y;
            } else {
                if (j0 < a.P) raw[0] = src[j0];
                if (j0 + 1 < a.P) raw[1] = src[j0 + 1];
            }
            double fsum = 0.0;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                if (!valid[k]) continue;
                float v = (float)raw[k];
                if (a.dark_avg) {
                    v = (float)__dsub_rn((double)v, davg[k]);
                    v = fmaxf(v, 0.0f);
                }
                if (!(v <= thresh[k])) {
                    v = (float)__dmul_rn((double)v, flat[k]);
                    const int pos = atomicAdd(&s_count, 1);
                    if (pos < kDfCap) {
                        s_pix[pos] = j0 + k;
                        s_t[pos] = t;
                        s_v[pos] = v;
                    } else a.summary[kSumOverflow] = 2;
                    fsum += (double)v;
                }
            }
            fsum = warp_sum(fsum);
            if ((tid & 31) == 0 && fsum != 0.0) atomicAdd(&s_fsum[fr - f_begin], fsum);
        }
        if (((fr - f_begin) % kDfFlushEvery) == kDfFlushEvery - 1) {
            // the last thread to arrive reads the final count, so the OR is the exact, CTA-uniform answer
            if (__syncthreads_or(s_count > kDfCap - kDfFlushEvery * kDfPixels)) flush();
        }
    }
    flush();
    for (int i = tid; i < f_end - f_begin; i += kDfThreads) {
        const int t = out_frame(a.first_raw + f_begin + i, a.rawblock, a.stride, a.F);
        if (t >= 0 && s_fsum[i] != 0.0) atomicAdd(a.frame_acc + t, s_fsum[i]);
    }
}

// ---- vectorised dense filter ------------------------------------------------------------
// dense_filter.cpp:150-174 again, arranged so that the 2 bytes per sample are all that moves.
// The whole test (dark subtraction in fp64, clamp, lld + sigma*std) is monotone in the raw
// int16 value, so it collapses, per pixel, into one int16 bound: a sample survives iff
// raw > bound.  k_dense_bounds finds that bound by bisection over the 65536 raw values with
// the very arithmetic of the reference.  The streaming loop keeps kDvGroup independent
// 16-byte loads (8 pixels x 8 frames) in flight per thread, reduces them to the per-pixel
// maximum over the group with the 16x2 SIMD max (VIMNMX3.S16x2), compares once against the
// bounds (VIMNMX.S16x2 with its two predicates), and only for a word whose maximum passes
// looks at the individual frames.  Candidates (pixel, frame slot, raw) go to private
// shared-memory slots -- no atomics in the loop; each warp settles and flushes its own
// candidates: the exact value is computed there, with the per-pixel constants fetched by all
// lanes at once, per-frame sums go to the CTA's shared accumulators, and the events are
// appended with one global atomic per flush and coalesced stores.
constexpr int kDvThreads = 128;
constexpr int kDvPix = 8;            // pixels per thread (one 16-byte load)
constexpr int kDvGroup = 8;          // frames in flight per thread
constexpr int kDvFrames = 64;        // frames per CTA
constexpr int kDvSlots = 12;         // private candidate slots per thread

__device__ __forceinline__ bool dense_sample(short raw, bool have_dark, double davg, float thresh, float &v)
{
    v = (float)raw;
    if (have_dark) {
        v = (float)__dsub_rn((double)v, davg);
        v = fmaxf(v, 0.0f);
    }
    return !(v <= thresh);
}

// bound[j]: survive iff raw > bound[j]; every[g] != 0 when a pixel of the 8-pixel group g lets
// even raw = -32768 through (the bound would be -32769), which makes every sample of the group
// a candidate for the exact test.
__global__ void k_dense_bounds(int P, const int *__restrict__ row_of_pixel, const double *__restrict__ dark_avg,
                               const double *__restrict__ dark_std, float lld, float sigma,
                               int16_t *__restrict__ bound, unsigned char *__restrict__ every)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g * kDvPix >= P) return;
    unsigned char ev = 0;
    for (int k = 0; k < kDvPix; k++) {
        const int j = g * kDvPix + k;
        if (j >= P) break;
        int b = 32767;
        if (row_of_pixel[j] >= 0) {
            const bool hd = dark_avg != nullptr;
            const double davg = hd ? dark_avg[j] : 0.0;
            const float thresh = hd ? (float)__dadd_rn((double)lld, __dmul_rn((double)sigma, dark_std[j])) : 0.0f;
            int lo = -32768, hi = 32768;  // smallest surviving raw value lies in [lo, hi]; hi = none
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                float v;
                if (dense_sample((short)mid, hd, davg, thresh, v)) hi = mid;
                else lo = mid + 1;
            }
            if (lo == -32768) {
                ev = 1;
                b = -32768;
            } else b = lo - 1;
        }
        bound[j] = (int16_t)b;
    }
    every[g] = ev;
}

// value of a sample that is known to survive (raw > bound[j], dense_filter.cpp:150-174): the
// bound already stands for the mask and the threshold
__device__ __forceinline__ float dense_value(const DenseArgs &a, int j, short raw)
{
    float v = (float)raw;
    if (a.dark_avg != nullptr) {
        v = (float)__dsub_rn((double)v, a.dark_avg[j]);
        v = fmaxf(v, 0.0f);
    }
    return (float)__dmul_rn((double)v, a.flat[j]);
}

// exact test and value of one sample (dense_filter.cpp:150-174)
__device__ __forceinline__ bool dense_exact(const DenseArgs &a, int j, short raw, float &v)
{
    if (a.row_of_pixel[j] < 0) return false;
    const bool have_dark = a.dark_avg != nullptr;
    const double davg = have_dark ? a.dark_avg[j] : 0.0;
    const float thresh = have_dark ? (float)__dadd_rn((double)a.lld, __dmul_rn((double)a.sigma, a.dark_std[j])) : 0.0f;
    if (!dense_sample(raw, have_dark, davg, thresh, v)) return false;
    v = (float)__dmul_rn((double)v, a.flat[j]);
    return true;
}

// a thread whose private slots are full (nearly every sample survives) settles the sample on
// the spot and appends it directly; kept out of line, the streaming loop has 64 call sites
__device__ __noinline__ void dense_settle_direct(const DenseArgs *ga, double *s_fsum, int j, unsigned raw16, int slot,
                                                 int raw_begin)
{
    // ga: the launch-invariant arguments in global memory (reading the kernel parameters from
    // an out-of-line function would cost every thread a stack copy of them)
    const DenseArgs &a = *ga;
    float v;
    if (!dense_exact(a, j, (short)raw16, v)) return;
    atomicAdd(&s_fsum[slot], (double)v);
    const unsigned long long gp = atomicAdd(a.counter, 1ull);
    if (gp < a.capacity) {
        a.out_idx[gp] = j;
        a.out_t[gp] = out_frame(raw_begin + slot, a.rawblock, a.stride, a.F);
        a.out_v[gp] = v;
    } else a.summary[kSumOverflow] = 2;
}

__global__ void __launch_bounds__(kDvThreads) k_dense_filter_vec(DenseArgs a, const DenseArgs *__restrict__ ga,
                                                                  const int16_t *__restrict__ bound,
                                                                  const unsigned char *__restrict__ every)
{
    __shared__ unsigned s_cand[kDvSlots][kDvThreads];   // private: (slot << 19 | pixel-in-group << 16 | raw)
    __shared__ int s_pix[kDvThreads * kDvSlots];        // per warp: settled events awaiting the copy-out
    __shared__ int s_slot[kDvThreads * kDvSlots];
    __shared__ float s_v[kDvThreads * kDvSlots];
    __shared__ double s_fsum[kDvFrames];
    __shared__ unsigned s_valid[kDvFrames / 32];
    __shared__ uint4 s_raw[kDvGroup][kDvThreads];        // private: the current group's samples of a thread with hits
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x * kDvThreads + tid;
    const int j0 = g * kDvPix;
    const bool active = j0 < a.P;
    const int f_begin = blockIdx.y * kDvFrames;
    const int f_end = min(a.nframes, f_begin + kDvFrames);
    for (int i = tid; i < kDvFrames; i += kDvThreads) s_fsum[i] = 0.0;
    uint4 bd = make_uint4(0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu);
    bool all_pass = false;
    if (active) {
        bd = *reinterpret_cast<const uint4 *>(bound + j0);
        all_pass = every[g] != 0;
    }
    int *w_pix = s_pix + warp * 32 * kDvSlots;
    int *w_slot = s_slot + warp * 32 * kDvSlots;
    float *w_v = s_v + warp * 32 * kDvSlots;
    int cnt = 0;
    if (tid < kDvFrames) {  // warps 0 and 1: which of the CTA's 64 frames exist and map to an output frame
        const bool ok = f_begin + tid < f_end && out_frame(a.first_raw + f_begin + tid, a.rawblock, a.stride, a.F) >= 0;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_valid[warp] = m;
    }
    __syncthreads();

    auto warp_flush = [&]() {  // all lanes of the warp
        __syncwarp();
        int x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        const int base = x - cnt;
        const int total = __shfl_sync(0xffffffffu, x, 31);
        for (int i = 0; i < cnt; i++) {
            const unsigned e = s_cand[i][tid];
            const int j = j0 + (int)((e >> 16) & 7u);
            const int slot = (int)(e >> 19);
            float v;
            bool ok = true;
            if (all_pass) ok = dense_exact(a, j, (short)(e & 0xffffu), v);  // bound -32768 is not exact
            else v = dense_value(a, j, (short)(e & 0xffffu));
            if (ok) {
                w_pix[base + i] = j;
                w_slot[base + i] = slot;
                w_v[base + i] = v;
                // per-frame sums stay in shared memory (fp64 compare-and-swap loops): 3e8 fire-and-forget
                // fp64 reductions in L2 instead were measured 4x slower for the whole kernel
                atomicAdd(&s_fsum[slot], (double)v);
            } else w_pix[base + i] = -1;
        }
        cnt = 0;
        __syncwarp();
        int npass = 0;
        for (int i0 = 0; i0 < total; i0 += 32) {
            const int i = i0 + lane;
            npass += __popc(__ballot_sync(0xffffffffu, i < total && w_pix[i] >= 0));
        }
        unsigned long long gbase = 0;
        if (lane == 0 && npass > 0) gbase = atomicAdd(a.counter, (unsigned long long)npass);
        gbase = __shfl_sync(0xffffffffu, gbase, 0);
        int run = 0;
        for (int i0 = 0; i0 < total; i0 += 32) {
            const int i = i0 + lane;
            const bool ok = i < total && w_pix[i] >= 0;
            const unsigned mk = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const unsigned long long pos = gbase + run + __popc(mk & ((1u << lane) - 1u));
                if (pos < a.capacity) {
                    a.out_idx[pos] = w_pix[i];
                    a.out_t[pos] = out_frame(a.first_raw + f_begin + w_slot[i], a.rawblock, a.stride, a.F);
                    a.out_v[pos] = w_v[i];
                } else a.summary[kSumOverflow] = 2;
            }
            run += __popc(mk);
        }
        __syncwarp();
    };

    auto candidate = [&](int k, unsigned raw16, int slot) {
        if (cnt < kDvSlots) {
            s_cand[cnt][tid] = ((unsigned)slot << 19) | ((unsigned)k << 16) | raw16;
            cnt++;
        } else dense_settle_direct(ga, s_fsum, j0 + k, raw16, slot, a.first_raw + f_begin);
    };

    for (int f0 = f_begin; f0 < f_end; f0 += kDvGroup) {
        // frames of this group that exist and map to an output frame (warp-uniform)
        const unsigned fmask = (s_valid[(f0 - f_begin) >> 5] >> ((f0 - f_begin) & 31)) & 0xffu;
        uint4 rw[kDvGroup];
        const uint4 *src = reinterpret_cast<const uint4 *>(a.frames + (int64_t)f0 * a.P + j0);
        const int fstride = a.P / kDvPix;  // frame pitch in 16-byte units
        if (active && fmask == 0xffu) {
#pragma unroll
            for (int u = 0; u < kDvGroup; u++) rw[u] = __ldcs(src + (int64_t)u * fstride);
        } else {
#pragma unroll
            for (int u = 0; u < kDvGroup; u++) {
                rw[u] = make_uint4(0x80008000u, 0x80008000u, 0x80008000u, 0x80008000u);
                if (active && ((fmask >> u) & 1u)) rw[u] = __ldcs(src + (int64_t)u * fstride);
            }
        }
        // per-pixel maximum over the group, one compare against the bounds per word
        bool hw[4];
        {
            unsigned mx[4];
            mx[0] = __vimax3_s16x2(rw[0].x, rw[1].x, rw[2].x);
            mx[1] = __vimax3_s16x2(rw[0].y, rw[1].y, rw[2].y);
            mx[2] = __vimax3_s16x2(rw[0].z, rw[1].z, rw[2].z);
            mx[3] = __vimax3_s16x2(rw[0].w, rw[1].w, rw[2].w);
            mx[0] = __vimax3_s16x2(mx[0], rw[3].x, rw[4].x);
            mx[1] = __vimax3_s16x2(mx[1], rw[3].y, rw[4].y);
            mx[2] = __vimax3_s16x2(mx[2], rw[3].z, rw[4].z);
            mx[3] = __vimax3_s16x2(mx[3], rw[3].w, rw[4].w);
            mx[0] = __vimax3_s16x2(mx[0], rw[5].x, rw[6].x);
            mx[1] = __vimax3_s16x2(mx[1], rw[5].y, rw[6].y);
            mx[2] = __vimax3_s16x2(mx[2], rw[5].z, rw[6].z);
            mx[3] = __vimax3_s16x2(mx[3], rw[5].w, rw[6].w);
            mx[0] = __vimax3_s16x2(mx[0], rw[7].x, rw[7].x);
            mx[1] = __vimax3_s16x2(mx[1], rw[7].y, rw[7].y);
            mx[2] = __vimax3_s16x2(mx[2], rw[7].z, rw[7].z);
            mx[3] = __vimax3_s16x2(mx[3], rw[7].w, rw[7].w);
            const unsigned bdw[4] = {bd.x, bd.y, bd.z, bd.w};
#pragma unroll
            for (int w = 0; w < 4; w++) {
                bool ph, pl;  // bound >= max: no frame of the group passes for that pixel
                (void)__vibmax_s16x2(bdw[w], mx[w], &ph, &pl);
                hw[w] = !(ph && pl) || all_pass;
            }
        }
        if (active && (hw[0] || hw[1] || hw[2] || hw[3])) {
            // which (frame, pixel) samples of the group pass: bit u*8 + k, frames 0-3 / 4-7 apart
            const unsigned bdw[4] = {bd.x, bd.y, bd.z, bd.w};
            unsigned mlo = 0u, mhi = 0u;
#pragma unroll
            for (int w = 0; w < 4; w++) {
                if (!hw[w]) continue;
#pragma unroll
                for (int u = 0; u < kDvGroup; u++) {
                    const unsigned word = w == 0 ? rw[u].x : w == 1 ? rw[u].y : w == 2 ? rw[u].z : rw[u].w;
                    bool ph, pl;
                    (void)__vibmax_s16x2(bdw[w], word, &ph, &pl);
                    const unsigned blo = 1u << ((u & 3) * 8 + 2 * w), bhi = blo << 1;
                    if (u < 4) {
                        if (!pl) mlo |= blo;
                        if (!ph) mlo |= bhi;
                    } else {
                        if (!pl) mhi |= blo;
                        if (!ph) mhi |= bhi;
                    }
                }
            }
            if (all_pass) mlo = mhi = 0xffffffffu;
            // frames that do not exist or map to no output frame drop out (fmask is warp-uniform)
            mlo &= ((fmask & 1u) ? 0xffu : 0u) | ((fmask & 2u) ? 0xff00u : 0u) | ((fmask & 4u) ? 0xff0000u : 0u) |
                   ((fmask & 8u) ? 0xff000000u : 0u);
            mhi &= ((fmask & 16u) ? 0xffu : 0u) | ((fmask & 32u) ? 0xff00u : 0u) | ((fmask & 64u) ? 0xff0000u : 0u) |
                   ((fmask & 128u) ? 0xff000000u : 0u);
            // the few survivors: re-read the sample (it has just been streamed through L2)
            if (mlo | mhi) {
                // the few survivors: park the group's samples in shared memory to pick them by index
#pragma unroll
                for (int u = 0; u < kDvGroup; u++) s_raw[u][tid] = rw[u];
                const unsigned short *mine = reinterpret_cast<const unsigned short *>(&s_raw[0][tid]);
                while (mlo | mhi) {
                    int b;
                    if (mlo) {
                        b = __ffs(mlo) - 1;
                        mlo &= mlo - 1u;
                    } else {
                        b = 32 + __ffs(mhi) - 1;
                        mhi &= mhi - 1u;
                    }
                    const int u = b >> 3, k = b & 7;
                    const unsigned raw16 = mine[u * (kDvThreads * kDvPix) + k];
                    candidate(k, raw16, f0 + u - f_begin);
                }
            }
        }
        if (__any_sync(0xffffffffu, cnt > kDvSlots - 4)) warp_flush();
    }
    warp_flush();
    __syncthreads();
    for (int i = tid; i < f_end - f_begin; i += kDvThreads) {
        const int t = out_frame(a.first_raw + f_begin + i, a.rawblock, a.stride, a.F);
        if (t >= 0 && s_fsum[i] != 0.0) atomicAdd(a.frame_acc + t, s_fsum[i]);
    }
}

// ------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------

static int max_dyn_smem(int device)
{
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    return v;
}

// chunk: -1 = whole ingest; k >= 0 = chunk k of a pipelined ingest (sums accumulate from k = 1 on)
template <int KIND>
static int run_store_build(xpcs_handle_s *h, IngestArgs &ia, int nblocks, bool dense, int chunk = -1)
{
    typedef typename WordT<KIND>::type W;
    // slice geometry from the histogram
    {
        LaunchScope ls(h, "k_slice_len");
        int threads = 256, warps = h->n_slices;
        k_slice_len<<<(warps * 32 + threads - 1) / threads, threads, 0, h->stream>>>(
            h->d_row_count.p, h->d_row_len.p, h->d_slice_len.p, h->d_slice_cur.p, h->n_slices, h->d_summary.p);
    }
    {
        LaunchScope ls(h, "k_slice_scan");
        k_slice_scan<<<2, 1024, 0, h->stream>>>(h->d_slice_len.p, h->d_slice_base.p, h->d_slice_cur.p,
                                                h->d_slice_rec.p, h->d_slice_end.p, h->n_slices, h->d_summary.p);
    }
    long long sum[kSumSlots];
    int rc = check_cuda(h, cudaMemcpyAsync(sum, h->d_summary.p, sizeof(sum), cudaMemcpyDeviceToHost, h->stream),
                        "summary D2H");
    if (rc) return rc;
    rc = check_cuda(h, cudaStreamSynchronize(h->stream), "ingest pass 1");
    if (rc) return rc;
    if (KIND == kPacked && sum[kSumBadCount]) return 1;  // caller retries with float values
    h->max_row = (int)sum[kSumMaxLen];
    h->store_words = sum[kSumWords];
    size_t need32 = (size_t)h->store_words * (sizeof(W) / 4);
    rc = ensure(h, h->d_store, chunk >= 0 ? need32 + need32 / 8 + 32 : need32 + 32, "event store");
    if (rc) return rc;
    ia.slice_base = h->d_slice_base.p;
    ia.store = h->d_store.p;
    // stream scatter + placement unless the frames do not fit the float record (or XPCS_SCATTER_DIRECT)
    const bool direct = (KIND == kFloat && h->prm.frames > (1 << 27)) || getenv("XPCS_SCATTER_DIRECT");
    // Short rows (a slice fits 24 KB): one lane per row, the whole slice in shared memory -- the
    // finalisation then takes the rows straight from the record streams.  Longer rows: one warp
    // per row; slices beyond its buffers are flagged and left to the lane-per-row kernel, which
    // then works in global memory.
    const int smem_cap = max_dyn_smem(h->device) - 1024;
    const size_t lane_bytes_all = (size_t)h->max_row * kSlice * sizeof(W);
    const bool use_warp = (lane_bytes_all > 24 * 1024 || getenv("XPCS_FIN_WARP")) && !getenv("XPCS_FIN_LANE");
    const bool fused_place = !direct && !use_warp && (long long)lane_bytes_all <= smem_cap && !getenv("XPCS_PLACE_KERNEL");
    if (nblocks > 0 && direct) {
        LaunchScope ls(h, dense ? "k_scatter_dense" : "k_scatter");
        if (dense) k_scatter<KIND, true><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
        else k_scatter<KIND, false><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
    } else if (nblocks > 0) {
        rc = ensure(h, h->d_rec, (size_t)sum[kSumEvents] + (chunk >= 0 ? (size_t)sum[kSumEvents] / 8 : 0) + 1, "event records");
        if (rc) return rc;
        ia.slice_end = h->d_slice_end.p;
        ia.rec = h->d_rec.p;
        {
            LaunchScope ls(h, dense ? "k_scatter_rec_dense" : "k_scatter_rec");
            if (dense) k_scatter_rec<KIND, true><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
            else k_scatter_rec<KIND, false><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
        }
        if (!fused_place) {
            LaunchScope ls(h, "k_place");
            k_place<KIND><<<h->n_slices, kPlaceWarps * 32, 0, h->stream>>>(h->d_rec.p, h->d_slice_rec.p,
                                                                          h->d_slice_base.p, h->d_store.p);
        }
    }
    // finalize
    const int F = h->prm.frames;
    const int windows = (F + h->prm.static_window - 1) / h->prm.static_window;
    if (chunk <= 0) {
        cudaMemsetAsync(h->d_part_total.p, 0, sizeof(double) * (size_t)h->S, h->stream);
        cudaMemsetAsync(h->d_part_partial.p, 0, sizeof(double) * (size_t)windows * h->S, h->stream);
    }
    FinalizeArgs fa{};
    fa.accumulate = chunk > 0 ? 1 : 0;
    fa.late_window = (h->prm.compat_flags & XPCS_COMPAT_LATE_WINDOW) ? 1 : 0;
    fa.store = h->d_store.p;
    fa.slice_base = h->d_slice_base.p;
    fa.slice_len = h->d_slice_len.p;
    fa.row_len = h->d_row_len.p;
    fa.sbin_of_row = h->d_sbin_of_row.p;
    fa.row_sum = h->d_row_sum.p;
    fa.part_total = h->d_part_total.p;
    fa.part_partial = h->d_part_partial.p;
    fa.frame_scale = nullptr;
    fa.summary = h->d_summary.p;
    fa.n_slices = h->n_slices;
    fa.S = h->S;
    fa.swindow = h->prm.static_window;
    fa.avg = h->prm.avg_frames;
    if (KIND == kFloat && h->prm.normalize_by_framesum) {
        rc = ensure(h, h->d_frame_scale, (size_t)F, "frame scale");
        if (rc) return rc;
        if (comm_active(h) && !h->frame_acc_reduced) {  // the scale needs the frame sums over ALL pixels (main.cpp:313-337)
            if ((rc = comm_allreduce_f64(h, h->d_frame_acc.p, (size_t)F))) return rc;
            h->frame_acc_reduced = true;
        }
        LaunchScope ls(h, "k_frame_scale");
        k_frame_scale<<<1, 256, 0, h->stream>>>(h->d_frame_acc.p, h->d_frame_scale.p, F, h->P);
        fa.frame_scale = h->d_frame_scale.p;
    }
    fa.rec = (fused_place && nblocks > 0) ? h->d_rec.p : nullptr;
    fa.slice_rec = h->d_slice_rec.p;
    fa.flagged = nullptr;
    if (use_warp && h->n_slices > 0) {
        rc = ensure(h, h->d_mt_fallback, (size_t)h->n_slices, "finalize flags");
        if (rc) return rc;
        cudaMemsetAsync(h->d_mt_fallback.p, 0, (size_t)h->n_slices, h->stream);
        int row_cap = std::min<int>(h->max_row, (int)(smem_cap / (kFwWarps * 2 * sizeof(W))));
        size_t wbytes = (size_t)kFwWarps * 2 * row_cap * sizeof(W);
        fa.flagged = h->d_mt_fallback.p;
        fa.row_cap = row_cap;
        fa.window = 16;
        if (const char *e = getenv("XPCS_FIN_WINDOW")) fa.window = std::max(1, atoi(e));
        rc = check_cuda(h, cudaFuncSetAttribute(k_finalize_warp<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)wbytes), "finalize_warp smem attr");
        if (rc) return rc;
        LaunchScope ls(h, "k_finalize_warp");
        k_finalize_warp<KIND><<<h->n_slices, kFwWarps * 32, wbytes, h->stream>>>(fa);
    }
    int smem_len = h->max_row;
    size_t bytes = (size_t)smem_len * kSlice * sizeof(W);
    if (use_warp) {  // only flagged (very long) slices arrive here: global-memory path
        smem_len = 0;
        bytes = 0;
    } else if ((long long)bytes > smem_cap) {
        smem_len = (int)(smem_cap / (kSlice * sizeof(W)));
        bytes = (size_t)smem_len * kSlice * sizeof(W);
    }
    fa.smem_len = smem_len;
    rc = check_cuda(h, cudaFuncSetAttribute(k_finalize<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)bytes), "finalize smem attr");
    if (rc) return rc;
    if (h->n_slices > 0 && (!use_warp || h->max_row > fa.row_cap)) {
        LaunchScope ls(h, "k_finalize");
        k_finalize<KIND><<<h->n_slices, 32, bytes, h->stream>>>(fa);
    }
    rc = check_cuda(h, cudaMemcpyAsync(sum, h->d_summary.p, sizeof(sum), cudaMemcpyDeviceToHost, h->stream),
                    "summary D2H");
    if (rc) return rc;
    rc = check_cuda(h, cudaStreamSynchronize(h->stream), "ingest pass 2");
    if (rc) return rc;
    if (KIND == kPacked && sum[kSumOverflow]) return 1;
    h->events_stored = sum[kSumEvents];
    h->max_count = std::max(h->max_count, KIND == kPacked ? std::max(1, (int)sum[kSumMaxCount]) : 0);
    return XPCS_OK;
}

int launch_ingest(xpcs_handle_s *h)
{
    const int F = h->prm.frames;
    const bool dense = h->dense_source;
    int rc;
    rc = ensure(h, h->d_summary, kSumSlots, "summary");
    if (rc) return rc;
    const int windows = (F + h->prm.static_window - 1) / h->prm.static_window;
    if ((rc = ensure(h, h->d_frame_acc, (size_t)F, "frame sums"))) return rc;
    if ((rc = ensure(h, h->d_row_sum, (size_t)h->R_pad, "row sums"))) return rc;
    if ((rc = ensure(h, h->d_part_total, (size_t)h->S, "partition sums"))) return rc;
    if ((rc = ensure(h, h->d_part_partial, (size_t)windows * h->S, "partition window sums"))) return rc;
    // (a pipelined ingest that had to be abandoned may have handed these to a chunk store)
    if ((rc = ensure(h, h->d_row_len, (size_t)h->R_pad, "row lengths"))) return rc;
    if ((rc = ensure(h, h->d_slice_base, (size_t)h->n_slices + 1, "slice offsets"))) return rc;

    IngestArgs ia{};
    int64_t E = h->E;
    if (dense) {
        unsigned long long cnt = 0;
        rc = check_cuda(h, cudaMemcpyAsync(&cnt, h->d_dense_counter.p, sizeof(cnt), cudaMemcpyDeviceToHost,
                                           h->stream), "dense counter");
        if (rc) return rc;
        if ((rc = check_cuda(h, cudaStreamSynchronize(h->stream), "dense filter"))) return rc;
        if (cnt > h->d_idx.n) return fail(h, XPCS_E_NOMEM, "dense event list overflow (%llu events)", cnt);
        E = (int64_t)cnt;
        h->E = E;
        ia.idx = h->d_idx.p;
        ia.evt = h->d_evt.p;
        ia.valf = h->d_valf.p;
    } else {
        ia.idx = h->ev_idx;
        ia.val = h->ev_val;
        ia.off = h->ev_off;
    }
    ia.nraw = h->raw_frames;
    ia.E = E;
    ia.row_of_pixel = h->d_row_of_pixel.p;
    ia.flat = h->d_flat.p;
    ia.row_count = h->d_row_count.p;
    ia.frame_acc = h->d_frame_acc.p;
    ia.summary = h->d_summary.p;
    ia.F = F;
    ia.stride = h->prm.stride_frames;
    ia.rawblock = h->prm.stride_frames > 1 ? h->prm.stride_frames : h->prm.avg_frames;
    if (h->prm.stride_frames > 1 && h->prm.avg_frames > 1) ia.rawblock = h->prm.stride_frames * h->prm.avg_frames;
    ia.P = h->P;
    const int nblocks = (int)((E + kEvPerBlock - 1) / kEvPerBlock);
    if (!dense) {
        if ((rc = ensure(h, h->d_block_first, (size_t)nblocks + 1, "block frames"))) return rc;
        ia.block_first = h->d_block_first.p;
        if (nblocks > 0) {
            LaunchScope ls(h, "k_block_frames");
            k_block_frames<<<(nblocks + 255) / 256, 256, 0, h->stream>>>(ia.off, ia.nraw, E, h->d_block_first.p,
                                                                       nblocks, 0);
        }
    }
    // exact integer path only for plain photon counts
    bool want_packed = !dense && h->flat_is_one && h->prm.avg_frames == 1 && !h->prm.normalize_by_framesum &&
                       F <= (1 << (32 - kCountBits));
    for (int attempt = 0; attempt < 2; attempt++) {
        const int kind = want_packed ? kPacked : kFloat;
        h->kind = kind;
        cudaMemsetAsync(h->d_summary.p, 0, sizeof(long long) * kSumSlots, h->stream);
        cudaMemsetAsync(h->d_row_count.p, 0, sizeof(int) * (size_t)h->R_pad, h->stream);
        if (!dense) cudaMemsetAsync(h->d_frame_acc.p, 0, sizeof(double) * (size_t)F, h->stream);
        if (nblocks > 0) {
            LaunchScope ls(h, dense ? "k_hist_dense" : "k_hist");
            if (dense) k_hist<kFloat, true><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
            else if (kind == kPacked) k_hist<kPacked, false><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
            else k_hist<kFloat, false><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
        }
        rc = (kind == kPacked) ? run_store_build<kPacked>(h, ia, nblocks, dense)
                               : run_store_build<kFloat>(h, ia, nblocks, dense);
        if (rc == 1 && kind == kPacked) {  // counts do not fit the packed word: redo as floats
            want_packed = false;
            continue;
        }
        if (rc) return rc;
        break;
    }
    return check_cuda(h, cudaGetLastError(), "ingest kernels");
}

// ---- pipelined sparse ingest ------------------------------------------------------------------
// Raw frames [f0, f1) -> chunk store h->chunk[h->pipe_chunks] (integer counts only: every sum the
// Filter stage produces is then an exact integer, so the chunk order cannot change a result).
int launch_ingest_chunk(xpcs_handle_s *h, int f0, int f1)
{
    const int F = h->prm.frames;
    const int k = h->pipe_chunks;
    if (k >= kMaxChunks) return 1;
    int rc;
    const int windows = (F + h->prm.static_window - 1) / h->prm.static_window;
    if ((rc = ensure(h, h->d_summary, kSumSlots, "summary"))) return rc;
    if ((rc = ensure(h, h->d_frame_acc, (size_t)F, "frame sums"))) return rc;
    if ((rc = ensure(h, h->d_row_sum, (size_t)h->R_pad, "row sums"))) return rc;
    if ((rc = ensure(h, h->d_part_total, (size_t)h->S, "partition sums"))) return rc;
    if ((rc = ensure(h, h->d_part_partial, (size_t)windows * h->S, "partition window sums"))) return rc;
    const int64_t e0 = h->frame_off_host[f0], e1 = h->frame_off_host[f1];
    IngestArgs ia{};
    ia.idx = h->d_idx.p;
    ia.val = h->d_val.p;
    ia.off = h->d_frame_off.p + f0;
    ia.ev_begin = e0;
    ia.frame_base = f0;
    ia.nraw = f1 - f0;
    ia.E = e1;
    ia.row_of_pixel = h->d_row_of_pixel.p;
    ia.flat = h->d_flat.p;
    ia.row_count = h->d_row_count.p;
    ia.frame_acc = h->d_frame_acc.p;
    ia.summary = h->d_summary.p;
    ia.F = F;
    ia.stride = 1;
    ia.rawblock = 1;
    ia.P = h->P;
    const int nblocks = (int)((e1 - (e0 & ~3LL) + kEvPerBlock - 1) / kEvPerBlock);
    if ((rc = ensure(h, h->d_block_first, (size_t)nblocks + 1, "block frames"))) return rc;
    ia.block_first = h->d_block_first.p;
    h->kind = kPacked;
    cudaMemsetAsync(h->d_summary.p, 0, sizeof(long long) * kSumSlots, h->stream);
    cudaMemsetAsync(h->d_row_count.p, 0, sizeof(int) * (size_t)h->R_pad, h->stream);
    if (k == 0) cudaMemsetAsync(h->d_frame_acc.p, 0, sizeof(double) * (size_t)F, h->stream);
    if (nblocks > 0) {
        {
            LaunchScope ls(h, "k_block_frames");
            k_block_frames<<<(nblocks + 255) / 256, 256, 0, h->stream>>>(ia.off, ia.nraw, ia.E, h->d_block_first.p,
                                                                       nblocks, ia.ev_begin);
        }
        LaunchScope ls(h, "k_hist");
        k_hist<kPacked, false><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
    }
    // the chunk's own buffers stand in for the handle's while it is built (same sizes from one ingest
    // to the next: no reallocation, which would synchronise the device and stall the copy stream)
    ChunkStore &c = h->chunk[k];
    std::swap(c.store, h->d_store);
    std::swap(c.slice_base, h->d_slice_base);
    std::swap(c.row_len, h->d_row_len);
    rc = ensure(h, h->d_row_len, (size_t)h->R_pad, "row lengths");
    if (!rc) rc = ensure(h, h->d_slice_base, (size_t)h->n_slices + 1, "slice offsets");
    if (!rc) rc = run_store_build<kPacked>(h, ia, nblocks, false, k);
    std::swap(c.store, h->d_store);
    std::swap(c.slice_base, h->d_slice_base);
    std::swap(c.row_len, h->d_row_len);
    if (rc) return rc;  // 1: a count does not fit the packed word
    h->pipe_events = (k == 0 ? 0 : h->pipe_events) + h->events_stored;
    h->pipe_chunks = k + 1;
    return check_cuda(h, cudaGetLastError(), "chunk ingest kernels");
}

// ---- stream mode (multitau_stream.cu) ----------------------------------------------------------
// One chunk of a frame stream: the events d_idx / d_val[0, nev) with the chunk-local frame offsets
// d_frame_off[0 .. nframes] are the raw frames [frame_base, frame_base + nframes).  Same passes as a chunk of the
// pipelined ingest; the store replaces chunk[0] (the previous chunk has been folded into the stream state by
// then -- same CUDA stream), frameSum / pixelSum / partition sums accumulate from the second chunk on.
// Returns 1 when a count does not fit the packed word.
int launch_ingest_stream_chunk(xpcs_handle_s *h, int frame_base, int nframes, int64_t nev, bool first)
{
    const int F = h->prm.frames;
    int rc;
    const int windows = (F + h->prm.static_window - 1) / h->prm.static_window;
    if ((rc = ensure(h, h->d_summary, kSumSlots, "summary"))) return rc;
    if ((rc = ensure(h, h->d_frame_acc, (size_t)F, "frame sums"))) return rc;
    if ((rc = ensure(h, h->d_row_sum, (size_t)h->R_pad, "row sums"))) return rc;
    if ((rc = ensure(h, h->d_part_total, (size_t)h->S, "partition sums"))) return rc;
    if ((rc = ensure(h, h->d_part_partial, (size_t)windows * h->S, "partition window sums"))) return rc;
    IngestArgs ia{};
    ia.idx = h->d_idx.p;
    ia.val = h->d_val.p;
    ia.off = h->d_frame_off.p;
    ia.ev_begin = 0;
    ia.frame_base = frame_base;
    ia.nraw = nframes;
    ia.E = nev;
    ia.row_of_pixel = h->d_row_of_pixel.p;
    ia.flat = h->d_flat.p;
    ia.row_count = h->d_row_count.p;
    ia.frame_acc = h->d_frame_acc.p;
    ia.summary = h->d_summary.p;
    ia.F = F;
    ia.stride = 1;
    ia.rawblock = 1;
    ia.P = h->P;
    const int nblocks = (int)((nev + kEvPerBlock - 1) / kEvPerBlock);
    if ((rc = ensure(h, h->d_block_first, (size_t)nblocks + 1, "block frames"))) return rc;
    ia.block_first = h->d_block_first.p;
    h->kind = kPacked;
    cudaMemsetAsync(h->d_summary.p, 0, sizeof(long long) * kSumSlots, h->stream);
    cudaMemsetAsync(h->d_row_count.p, 0, sizeof(int) * (size_t)h->R_pad, h->stream);
    if (first) {
        cudaMemsetAsync(h->d_frame_acc.p, 0, sizeof(double) * (size_t)F, h->stream);
        cudaMemsetAsync(h->d_row_sum.p, 0, sizeof(double) * (size_t)h->R_pad, h->stream);
    }
    if (nblocks > 0) {
        {
            LaunchScope ls(h, "k_block_frames");
            k_block_frames<<<(nblocks + 255) / 256, 256, 0, h->stream>>>(ia.off, ia.nraw, ia.E, h->d_block_first.p, nblocks, 0);
        }
        LaunchScope ls(h, "k_hist");
        k_hist<kPacked, false><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
    }
    ChunkStore &c = h->chunk[0];
    std::swap(c.store, h->d_store);
    std::swap(c.slice_base, h->d_slice_base);
    std::swap(c.row_len, h->d_row_len);
    rc = ensure(h, h->d_row_len, (size_t)h->R_pad, "row lengths");
    if (!rc) rc = ensure(h, h->d_slice_base, (size_t)h->n_slices + 1, "slice offsets");
    if (!rc) rc = run_store_build<kPacked>(h, ia, nblocks, false, first ? 0 : 1);
    std::swap(c.store, h->d_store);
    std::swap(c.slice_base, h->d_slice_base);
    std::swap(c.row_len, h->d_row_len);
    if (rc) return rc;
    return check_cuda(h, cudaGetLastError(), "stream chunk ingest kernels");
}

struct ConcatArgs {
    const uint32_t *store[kMaxChunks];
    const int64_t *slice_base[kMaxChunks];
    const int *row_len[kMaxChunks];
    int K;
};

__global__ void k_chunk_rows(ConcatArgs c, int *__restrict__ row_count, int R_pad)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R_pad) return;
    int n = 0;
    for (int k = 0; k < c.K; k++) n += c.row_len[k][r];
    row_count[r] = n;
}

// One warp per slice, one lane per row: the chunk rows (sorted, chunks in frame order) one after
// the other; the slice is assembled in shared memory and leaves as full 128-byte lines.
__global__ void __launch_bounds__(32) k_concat(ConcatArgs c, uint32_t *__restrict__ store,
                                               const int64_t *__restrict__ slice_base,
                                               const int *__restrict__ slice_len, const int *__restrict__ row_len,
                                               int smem_len)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = blockIdx.x, lane = threadIdx.x;
    const int r = s * kSlice + lane;
    const int len = slice_len[s];
    if (len == 0) return;
    const int n = row_len[r];
    uint32_t *g = store + slice_base[s] + lane;
    const bool in_smem = len <= smem_len;
    uint32_t *tile = reinterpret_cast<uint32_t *>(smem_raw) + lane;
    int off = 0;
    for (int k = 0; k < c.K; k++) {
        const int nk = c.row_len[k][r];
        const int lenk = __reduce_max_sync(0xffffffffu, nk);
        const uint32_t *src = c.store[k] + c.slice_base[k][s] + lane;
        for (int j = 0; j < lenk; j++)
            if (j < nk) {
                const uint32_t w = src[(int64_t)j * kSlice];
                if (in_smem) tile[(off + j) * kSlice] = w;
                else g[(int64_t)(off + j) * kSlice] = w;
            }
        off += nk;
    }
    if (in_smem) {
        __syncwarp();
        for (int j = 0; j < len; j++)
            if (j < n) g[(int64_t)j * kSlice] = tile[j * kSlice];
    }
}

int launch_ingest_concat(xpcs_handle_s *h)
{
    int rc;
    ConcatArgs c{};
    c.K = h->pipe_chunks;
    for (int k = 0; k < c.K; k++) {
        c.store[k] = h->chunk[k].store.p;
        c.slice_base[k] = h->chunk[k].slice_base.p;
        c.row_len[k] = h->chunk[k].row_len.p;
    }
    if ((rc = ensure(h, h->d_row_len, (size_t)h->R_pad, "row lengths"))) return rc;
    if ((rc = ensure(h, h->d_slice_base, (size_t)h->n_slices + 1, "slice offsets"))) return rc;
    cudaMemsetAsync(h->d_summary.p, 0, sizeof(long long) * kSumSlots, h->stream);
    {
        LaunchScope ls(h, "k_chunk_rows");
        k_chunk_rows<<<(h->R_pad + 255) / 256, 256, 0, h->stream>>>(c, h->d_row_count.p, h->R_pad);
    }
    {
        LaunchScope ls(h, "k_slice_len");
        int threads = 256, warps = h->n_slices;
        k_slice_len<<<(warps * 32 + threads - 1) / threads, threads, 0, h->stream>>>(
            h->d_row_count.p, h->d_row_len.p, h->d_slice_len.p, h->d_slice_cur.p, h->n_slices, h->d_summary.p);
    }
    {
        LaunchScope ls(h, "k_slice_scan");
        k_slice_scan<<<2, 1024, 0, h->stream>>>(h->d_slice_len.p, h->d_slice_base.p, h->d_slice_cur.p,
                                                h->d_slice_rec.p, h->d_slice_end.p, h->n_slices, h->d_summary.p);
    }
    long long sum[kSumSlots];
    rc = check_cuda(h, cudaMemcpyAsync(sum, h->d_summary.p, sizeof(sum), cudaMemcpyDeviceToHost, h->stream), "summary D2H");
    if (rc) return rc;
    if ((rc = check_cuda(h, cudaStreamSynchronize(h->stream), "chunk totals"))) return rc;
    h->max_row = (int)sum[kSumMaxLen];
    h->store_words = sum[kSumWords];
    h->events_stored = h->pipe_events;  // as the one-pass ingest counts them: before duplicates merge
    h->kind = kPacked;
    if ((rc = ensure(h, h->d_store, (size_t)h->store_words + 32, "event store"))) return rc;
    const int smem_cap = max_dyn_smem(h->device) - 1024;
    int smem_len = h->max_row;
    size_t bytes = (size_t)smem_len * kSlice * sizeof(uint32_t);
    if ((long long)bytes > smem_cap) {
        smem_len = 0;  // very long rows: straight to global memory
        bytes = 0;
    }
    if ((rc = check_cuda(h, cudaFuncSetAttribute(k_concat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                         "concat smem attr")))
        return rc;
    if (h->n_slices > 0) {
        LaunchScope ls(h, "k_concat");
        k_concat<<<h->n_slices, 32, bytes, h->stream>>>(c, h->d_store.p, h->d_slice_base.p, h->d_slice_len.p,
                                                         h->d_row_len.p, smem_len);
    }
    return check_cuda(h, cudaGetLastError(), "concat kernels");
}

int launch_dark(xpcs_handle_s *h, const int16_t *d_frames, int n)
{
    int rc;
    if ((rc = ensure(h, h->d_dark_avg, (size_t)h->P, "dark avg"))) return rc;
    if ((rc = ensure(h, h->d_dark_std, (size_t)h->P, "dark std"))) return rc;
    {
        LaunchScope ls(h, "k_dark");
        k_dark<<<(h->P + 255) / 256, 256, 0, h->stream>>>(d_frames, n, h->P, h->d_flat.p, h->d_dark_avg.p,
                                                         h->d_dark_std.p);
    }
    return check_cuda(h, cudaGetLastError(), "k_dark");
}

int launch_dense_filter(xpcs_handle_s *h, const int16_t *d_frames, int first_raw, int nframes)
{
    DenseArgs a{};
    a.frames = d_frames;
    a.first_raw = first_raw;
    a.nframes = nframes;
    a.P = h->P;
    a.F = h->prm.frames;
    a.stride = h->prm.stride_frames;
    a.rawblock = h->prm.stride_frames > 1 ? h->prm.stride_frames : h->prm.avg_frames;
    if (h->prm.stride_frames > 1 && h->prm.avg_frames > 1) a.rawblock = h->prm.stride_frames * h->prm.avg_frames;
    a.row_of_pixel = h->d_row_of_pixel.p;
    a.flat = h->d_flat.p;
    a.dark_avg = h->have_dark ? h->d_dark_avg.p : nullptr;
    a.dark_std = h->have_dark ? h->d_dark_std.p : nullptr;
    a.lld = h->prm.lld;
    a.sigma = h->prm.sigma;
    a.out_idx = h->d_idx.p;
    a.out_t = h->d_evt.p;
    a.out_v = h->d_valf.p;
    a.counter = h->d_dense_counter.p;
    a.capacity = h->d_idx.n;
    a.frame_acc = h->d_frame_acc.p;
    a.summary = h->d_summary.p;
    // vector path: whole 16-byte groups of pixels, aligned frames
    const bool vec = (h->P % kDvPix) == 0 && (reinterpret_cast<uintptr_t>(d_frames) & 15u) == 0 &&
                     !(h->prm.compat_flags & XPCS_FLAG_SCALAR_DENSE);
    if (vec) {
        if (!h->dense_bounds_ready) {
            int rc;
            if ((rc = ensure(h, h->d_dense_bound, (size_t)h->P, "dense bounds"))) return rc;
            if ((rc = ensure(h, h->d_dense_every, (size_t)(h->P / kDvPix), "dense bounds"))) return rc;
            LaunchScope ls(h, "k_dense_bounds");
            const int groups = h->P / kDvPix;
            k_dense_bounds<<<(groups + 127) / 128, 128, 0, h->stream>>>(h->P, a.row_of_pixel, a.dark_avg, a.dark_std, a.lld,
                                                                        a.sigma, h->d_dense_bound.p, h->d_dense_every.p);
            h->dense_bounds_ready = true;
        }
        // the launch-invariant arguments once per ingest in global memory, for the out-of-line path
        DenseArgs *ga = reinterpret_cast<DenseArgs *>(h->d_dense_args.p);
        if (!h->dense_args_ready) {
            int rc;
            if ((rc = ensure(h, h->d_dense_args, sizeof(DenseArgs), "dense args"))) return rc;
            ga = reinterpret_cast<DenseArgs *>(h->d_dense_args.p);
            rc = check_cuda(h, cudaMemcpyAsync(ga, &a, sizeof(DenseArgs), cudaMemcpyHostToDevice, h->stream), "dense args");
            if (rc) return rc;
            h->dense_args_ready = true;
        }
        dim3 grid((h->P / kDvPix + kDvThreads - 1) / kDvThreads, (nframes + kDvFrames - 1) / kDvFrames);
        LaunchScope ls(h, "k_dense_filter");
        k_dense_filter_vec<<<grid, kDvThreads, 0, h->stream>>>(a, ga, h->d_dense_bound.p, h->d_dense_every.p);
    } else {
        dim3 grid((h->P + kDfPixels - 1) / kDfPixels, (nframes + kDfFrames - 1) / kDfFrames);
        LaunchScope ls(h, "k_dense_filter");
        k_dense_filter<<<grid, kDfThreads, 0, h->stream>>>(a);
    }
    return check_cuda(h, cudaGetLastError(), "k_dense_filter");
}

}  // namespace xpcs
