// multitau_stream.cu -- ONLINE multi-tau: the frame stream is correlated chunk by chunk, no row is ever complete
// on the device (C5 at high occupancy: 2e11 events do not fit 8 x 180 GB).
//
// The reference is offline: main.cpp:263-268 reads every frame, Corr::multiTau2 (corr.cpp:315-431) then bins the
// levels of a complete row in place (corr.cpp:349-390).  Here xpcs_stream_push_sparse hands over 2^k frames at a
// time; the Filter stage builds the pixel-major store of THAT chunk exactly as the pipelined ingest does
// (ingest.cu: histogram -> slice streams -> finalise, Filter sums accumulating), k_stream_chunk folds it into a
// per-row state -- integer G2 numerators, per-level totals, the last and the first 2 dpl bins of every level, the
// open bins of the levels above k -- and the chunk store is overwritten by the next one.  k_stream_finish turns the
// state into G2 / IP / IF with the one IEEE division per output of the resident kernels, so the rest of the path
// (xpcs_normalize, the result getters, the NCCL reduction of the partials) runs unchanged.  The arithmetic lives
// in multitau_stream_core.h, which is also compiled for the host and checked against the oracle on the CPU.
//
// One warp per row, 4-8 consecutive rows of a slice per CTA (their words share 32-byte sectors of the store).  Per
// warp in shared memory: the chunk expanded into a dense bin array ((2 dpl + 2^k) words, halved in place from level to
// level), the chunk-local frames of the row's events (16 bits each) and a copy of the row's state, which comes in and
// goes out with full-width accesses once per row and chunk.  A level is walked by its bins (four per lane and step,
// the 2 dpl bins in front of them from five 16-byte loads, 16 / 8 delays at once) or, while the row has fewer than
// half as many events as the level has bins, by the events; rows without an event in the chunk only shift their
// tails.  Measured on B200 (profiles/r04_*): 1.50 ms per chunk of 205 684 rows x 2048 frames at 5 %
// occupancy (1.37e8 rows x chunks per second, the same at any occupancy for the bin walk), 3 874 warp instructions per
// row and chunk before the shared-memory staging of the state; a 5 % job streams in 24.5 ms against 23.3 ms resident,
// and 2048 x 2048 pixels at 5 % run at 27 300 frames/s whatever the frame count (one PCIe link delivers 4.4e4
// frames/s of that detector: on one GPU the device, not the link, is the limit).
#include <algorithm>
#include <cstdlib>

#include "internal.h"

#include "multitau_stream_core.h"

namespace xpcs {

struct StArgs {
    const uint32_t *store;
    const int64_t *slice_base;
    const int *row_len;      // nullptr: the chunk has no event at all
    uint32_t *state;
    int R, c, stride_words, x_words;
    st::StSched s;
};

constexpr int kStMaxWarps = 8;

template <int DPL>
__global__ void __launch_bounds__(kStMaxWarps * 32) k_stream_chunk(StArgs a, int warps)
{
    extern __shared__ __align__(16) uint32_t st_smem[];
    const int warp = threadIdx.x >> 5;
    const int r = blockIdx.x * warps + warp;
    if (r >= a.R) return;  // (no CTA-wide barrier anywhere below)
    const int n = a.row_len ? a.row_len[r] : 0;
    const uint32_t *ev = a.store + (n > 0 ? a.slice_base[r >> 5] : 0) + (r & 31);
    st::row_chunk<DPL>(a.s, a.c, ev, n, st_smem + (size_t)warp * a.x_words, a.state + (size_t)r * a.stride_words);
}

template <int DPL>
__global__ void k_stream_finish(const uint32_t *__restrict__ state, int stride_words, st::StSched s, float *__restrict__ G2,
                                float *__restrict__ IP, float *__restrict__ IF, int R, int R_pad)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const uint32_t *row = state + (size_t)r * stride_words;
    for (int ti = 0; ti < s.T; ti++) {
        float g, p, f;
        st::row_result<DPL>(s, row, ti, g, p, f);
        const size_t o = (size_t)ti * R_pad + r;
        G2[o] = g;
        IP[o] = p;
        IF[o] = f;
    }
}

static st::StSched stream_sched(const xpcs_handle_s *h)
{
    st::StSched s{};
    const Sched &sc = h->sched;
    s.F = sc.frames;
    s.T = h->T;
    s.cnt0 = sc.count[0];
    s.lastl = 0;
    s.cnt_last = 0;
    for (int l = 1; l < sc.n_levels; l++)
        if (sc.count[l] > 0) {
            s.lastl = l;
            s.cnt_last = sc.count[l];
        }
    s.k = h->stream_k;
    s.ev_num = 1;
    if (const char *e = getenv("XPCS_ST_EVENTS")) s.ev_num = std::max(0, atoi(e));  // diagnostics: 0 = every level by its bins
    return s;
}

static int stream_state_words(const xpcs_handle_s *h, const st::StSched &s)
{
    return h->prm.delays_per_level == 8 ? st::Layout<8>::words(s) : st::Layout<4>::words(s);
}

int stream_check(xpcs_handle_s *h, int chunk_frames)
{
    const int dpl = h->prm.delays_per_level;
    if (dpl != 4 && dpl != 8) return fail(h, XPCS_E_ARG, "stream mode is built for delays_per_level 4 and 8 (got %d)", dpl);
    if (chunk_frames < 64 || chunk_frames > 8192 || (chunk_frames & (chunk_frames - 1)))
        return fail(h, XPCS_E_ARG, "stream mode: chunk_frames must be a power of two in [64, 8192] (got %d)", chunk_frames);
    if (h->prm.compat_flags & XPCS_COMPAT_STALE_TAIL)
        return fail(h, XPCS_E_ARG, "stream mode computes the exact sums: XPCS_COMPAT_STALE_TAIL needs the complete rows "
                                   "(SURVEY.md A.4); clear the flag");
    if (!h->flat_is_one || h->prm.avg_frames != 1 || h->prm.stride_frames != 1 || h->prm.normalize_by_framesum)
        return fail(h, XPCS_E_ARG, "stream mode takes plain photon counts: no flat field, stride, averaging or frame-sum normalisation");
    if (h->prm.frames > (1 << (32 - kCountBits))) return fail(h, XPCS_E_ARG, "stream mode: more than 2^20 frames");
    // the schedule must be the regular one the core assumes: every level between 1 and the last has dpl delays
    const Sched &sc = h->sched;
    int lastl = 0;
    for (int l = 1; l < sc.n_levels; l++)
        if (sc.count[l] > 0) lastl = l;
    if (sc.lo[0] != 1 || sc.count[0] > 2 * dpl) return fail(h, XPCS_E_ARG, "stream mode: irregular delay schedule");
    for (int l = 1; l <= lastl; l++)
        if (sc.lo[l] != dpl + 1 || sc.count[l] > dpl || (l < lastl && sc.count[l] != dpl) || sc.first[l] != sc.count[0] + (l - 1) * dpl)
            return fail(h, XPCS_E_ARG, "stream mode: irregular delay schedule");
    return XPCS_OK;
}

int launch_stream_begin(xpcs_handle_s *h)
{
    const st::StSched s = stream_sched(h);
    const size_t words = (size_t)stream_state_words(h, s) * (size_t)std::max(h->R_pad, 1);
    int rc = ensure(h, h->d_stream_state, words, "stream state");
    if (rc) return rc;
    return check_cuda(h, cudaMemsetAsync(h->d_stream_state.p, 0, sizeof(uint32_t) * words, h->stream), "stream state clear");
}

int launch_stream_chunk(xpcs_handle_s *h, int c, bool empty)
{
    if (h->R == 0) return XPCS_OK;
    StArgs a{};
    a.s = stream_sched(h);
    const ChunkStore &cs = h->chunk[0];
    a.store = cs.store.p;
    a.slice_base = cs.slice_base.p;
    a.row_len = empty ? nullptr : cs.row_len.p;
    a.state = h->d_stream_state.p;
    a.R = h->R;
    a.c = c;
    a.stride_words = stream_state_words(h, a.s);
    const int dpl = h->prm.delays_per_level;
    a.x_words = dpl == 8 ? st::Layout<8>::scratch_words(a.s) : st::Layout<4>::scratch_words(a.s);  // per warp: bins, event frames, state copy
    int smem_cap = 0;
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    const size_t per_warp = sizeof(uint32_t) * (size_t)a.x_words;
    int warps = (int)std::min<size_t>(kStMaxWarps, (size_t)(smem_cap - 1024) / per_warp);
    {   // shared memory decides how many warps an SM holds: the CTA size (4..8 warps) that leaves the least of it unused
        int smem_sm = 0;
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device);
        int best = 0;
        for (int w = warps; w >= std::min(warps, 4); w--) {
            const int resident = w * (int)((size_t)smem_sm / (per_warp * w + 1024));
            if (resident > best) {
                best = resident;
                warps = w;
            }
        }
    }
    if (const char *e = getenv("XPCS_ST_WARPS")) warps = std::max(1, std::min(warps, atoi(e)));  // diagnostics
    if (warps < 1) return fail(h, XPCS_E_ARG, "stream mode: chunk too long for shared memory");
    const size_t bytes = per_warp * warps;
    const int grid = (h->R + warps - 1) / warps;
    int rc;
    LaunchScope ls(h, "k_stream_chunk");
    if (dpl == 8) {
        if ((rc = check_cuda(h, cudaFuncSetAttribute(k_stream_chunk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "stream smem attr"))) return rc;
        k_stream_chunk<8><<<grid, warps * 32, bytes, h->stream>>>(a, warps);
    } else {
        if ((rc = check_cuda(h, cudaFuncSetAttribute(k_stream_chunk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "stream smem attr"))) return rc;
        k_stream_chunk<4><<<grid, warps * 32, bytes, h->stream>>>(a, warps);
    }
    return check_cuda(h, cudaGetLastError(), "k_stream_chunk");
}

int launch_stream_finish(xpcs_handle_s *h)
{
    int rc;
    const size_t n = (size_t)h->T * h->R_pad;
    if ((rc = ensure(h, h->d_G2, n, "G2"))) return rc;
    if ((rc = ensure(h, h->d_IP, n, "IP"))) return rc;
    if ((rc = ensure(h, h->d_IF, n, "IF"))) return rc;
    cudaMemsetAsync(h->d_G2.p, 0, sizeof(float) * n, h->stream);
    cudaMemsetAsync(h->d_IP.p, 0, sizeof(float) * n, h->stream);
    cudaMemsetAsync(h->d_IF.p, 0, sizeof(float) * n, h->stream);
    if (h->R > 0) {
        const st::StSched s = stream_sched(h);
        const int stride = stream_state_words(h, s);
        LaunchScope ls(h, "k_stream_finish");
        if (h->prm.delays_per_level == 8)
            k_stream_finish<8><<<(h->R + 127) / 128, 128, 0, h->stream>>>(h->d_stream_state.p, stride, s, h->d_G2.p, h->d_IP.p, h->d_IF.p, h->R, h->R_pad);
        else
            k_stream_finish<4><<<(h->R + 127) / 128, 128, 0, h->stream>>>(h->d_stream_state.p, stride, s, h->d_G2.p, h->d_IP.p, h->d_IF.p, h->R, h->R_pad);
    }
    return check_cuda(h, cudaGetLastError(), "k_stream_finish");
}

}  // namespace xpcs
