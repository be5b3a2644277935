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

#include "internal.h"

namespace xpcs {

// ------------------------------------------------------------------------------------
// frame lookup helpers (sparse IMM source: events are concatenated per raw frame)
// ------------------------------------------------------------------------------------

// first raw frame of every block of kEvPerBlock events: largest f with off[f] <= e
__global__ void k_block_frames(const int64_t *__restrict__ off, int nraw, int64_t E,
                               int *__restrict__ first, int nblocks)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    int64_t e = (int64_t)b * kEvPerBlock;
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
};

// summary slots
enum { kSumBadCount = 0, kSumMaxLen = 1, kSumEvents = 2, kSumOverflow = 3, kSumWords = 4, kSumSlots = 8 };

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
    const int64_t base = (int64_t)blockIdx.x * kEvPerBlock;
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
            if (j >= nv) { t[j] = -1; continue; }
            if (!DENSE_SRC) {
                const int64_t e = e0 + j;
                while (f + 1 < a.nraw && e >= __ldg(a.off + f + 1)) f++;
                t[j] = out_frame(f, a.rawblock, a.stride, a.F);
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
                            int *__restrict__ slice_len, int n_slices, long long *summary)
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
        atomicMax(summary + kSumMaxLen, (long long)m);
        atomicAdd((unsigned long long *)summary + kSumEvents, (unsigned long long)tot);
    }
}

// exclusive scan of slice_len*32 over all slices (single CTA; a few 10^4..10^5 entries)
__global__ void __launch_bounds__(1024) k_slice_scan(const int *__restrict__ slice_len,
                                                     int64_t *__restrict__ slice_base,
                                                     int n_slices, long long *summary)
{
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const int per = (n_slices + 1023) / 1024;
    const int a = tid * per, b = min(n_slices, a + per);
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
    const int64_t base = (int64_t)blockIdx.x * kEvPerBlock;
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
            if (j >= nv) continue;
            int tj;
            if (DENSE_SRC) tj = t[j];
            else {
                const int64_t e = e0 + j;
                while (f + 1 < a.nraw && e >= __ldg(a.off + f + 1)) f++;
                tj = out_frame(f, a.rawblock, a.stride, a.F);
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
};

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

template <int KIND>
__global__ void __launch_bounds__(32) k_finalize(FinalizeArgs a)
{
    typedef typename WordT<KIND>::type W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = blockIdx.x;
    const int lane = threadIdx.x;
    const int r = s * kSlice + lane;
    const int len = a.slice_len[s];
    if (len == 0) {
        a.row_sum[r] = 0.0;
        return;
    }
    W *g = reinterpret_cast<W *>(a.store) + a.slice_base[s] + lane;
    const int n = a.row_len[r];
    const bool in_smem = len <= a.smem_len;
    W *col;
    if (in_smem) {
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
    // insertion sort by frame (the key sits in the high bits of the word)
    for (int i = 1; i < n; i++) {
        W w = col[(int64_t)i * kSlice];
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
    {
        float fsum = 0.0f;
        long long isum = 0;
        int win = -1;
        double wacc = 0.0;
        for (int i = 0; i < m; i++) {
            int t;
            double v;
            if (KIND == kPacked) {
                uint32_t w = (uint32_t)col[(int64_t)i * kSlice];
                t = (int)(w >> kCountBits);
                int c = (int)(w & ((1u << kCountBits) - 1));
                isum += c;
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
            int wi = t / a.swindow;
            if (wi != win) {
                if (win >= 0 && sb >= 0) atomicAdd(a.part_partial + (int64_t)win * a.S + sb, wacc);
                win = wi;
                wacc = 0.0;
            }
            wacc += v;
        }
        if (win >= 0 && sb >= 0) atomicAdd(a.part_partial + (int64_t)win * a.S + sb, wacc);
        total = (KIND == kPacked) ? (double)isum : (double)fsum;
    }
    a.row_sum[r] = total;
    if (sb >= 0 && m > 0) atomicAdd(a.part_total + sb, total);
    a.row_len[r] = m;
    if (in_smem) {
        __syncwarp();
        for (int j = 0; j < len; j++)
            if (j < m) g[(int64_t)j * kSlice] = col[j * kSlice];
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
            __syncthreads();
            if (s_count > kDfCap - kDfFlushEvery * kDfPixels) flush();
        }
    }
    flush();
    for (int i = tid; i < f_end - f_begin; i += kDfThreads) {
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

template <int KIND>
static int run_store_build(xpcs_handle_s *h, IngestArgs &ia, int nblocks, bool dense)
{
    typedef typename WordT<KIND>::type W;
    // slice geometry from the histogram
    {
        LaunchScope ls(h, "k_slice_len");
        int threads = 256, warps = h->n_slices;
        k_slice_len<<<(warps * 32 + threads - 1) / threads, threads, 0, h->stream>>>(
            h->d_row_count.p, h->d_row_len.p, h->d_slice_len.p, h->n_slices, h->d_summary.p);
    }
    {
        LaunchScope ls(h, "k_slice_scan");
        k_slice_scan<<<1, 1024, 0, h->stream>>>(h->d_slice_len.p, h->d_slice_base.p, h->n_slices,
                                                h->d_summary.p);
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
    rc = ensure(h, h->d_store, need32 + 32, "event store");
    if (rc) return rc;
    ia.slice_base = h->d_slice_base.p;
    ia.store = h->d_store.p;
    if (nblocks > 0) {
        LaunchScope ls(h, dense ? "k_scatter_dense" : "k_scatter");
        if (dense) k_scatter<KIND, true><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
        else k_scatter<KIND, false><<<nblocks, kIngestThreads, 0, h->stream>>>(ia);
    }
    // finalize
    const int F = h->prm.frames;
    const int windows = (F + h->prm.static_window - 1) / h->prm.static_window;
    cudaMemsetAsync(h->d_part_total.p, 0, sizeof(double) * (size_t)h->S, h->stream);
    cudaMemsetAsync(h->d_part_partial.p, 0, sizeof(double) * (size_t)windows * h->S, h->stream);
    FinalizeArgs fa{};
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
        LaunchScope ls(h, "k_frame_scale");
        k_frame_scale<<<1, 256, 0, h->stream>>>(h->d_frame_acc.p, h->d_frame_scale.p, F, h->P);
        fa.frame_scale = h->d_frame_scale.p;
    }
    const int smem_cap = max_dyn_smem(h->device) - 1024;
    int smem_len = h->max_row;
    size_t bytes = (size_t)smem_len * kSlice * sizeof(W);
    if ((long long)bytes > smem_cap) {
        smem_len = (int)(smem_cap / (kSlice * sizeof(W)));
        // keep several warps per SM when only a few slices are long
        bytes = (size_t)smem_len * kSlice * sizeof(W);
    }
    fa.smem_len = smem_len;
    rc = check_cuda(h, cudaFuncSetAttribute(k_finalize<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)bytes), "finalize smem attr");
    if (rc) return rc;
    if (h->n_slices > 0) {
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
                                                                       nblocks);
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
    dim3 grid((h->P + kDfPixels - 1) / kDfPixels, (nframes + kDfFrames - 1) / kDfFrames);
    {
        LaunchScope ls(h, "k_dense_filter");
        k_dense_filter<<<grid, kDfThreads, 0, h->stream>>>(a);
    }
    return check_cuda(h, cudaGetLastError(), "k_dense_filter");
}

}  // namespace xpcs
