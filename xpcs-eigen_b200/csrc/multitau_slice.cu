// multitau_slice.cu -- multi-tau correlator for short rows: lane = pixel row, warps = tasks.
//
// Replaces Corr::multiTau2 (reference corr.cpp:315-431) for the packed store (word = frame << 12 | count)
// when the rows of a slice fit a small shared-memory tile (typical sparse XPCS: ~10^2 events per row).
// One CTA owns one slice of 32 rows.  The tile is copied into shared memory as it lies in the store
// (one contiguous block, 16-byte loads; bank == lane afterwards), and the warps then run DIFFERENT
// jobs on the same 32 rows, every lane on its own row with plain sequential code
// (multitau_slice_core.h -- the same source is compiled for the host and checked against the oracle
// on the CPU):
//     pair warps     G2 of the sparse levels: one flattened walk over the event pairs of a row,
//                    events dealt round-robin to the pair warps, shared-memory atomics on H[slot][lane]
//     IF warp        forward walk: running sum of the counts at the ascending thresholds t' << l
//     IP warp        backward walk at the descending thresholds (L_l - t') << l
//     dense warps    G2 of the dense levels: (level, bin range) pieces; the bins are formed on the fly
//                    under a sliding register window, no bin array
// IP and IF leave straight from registers, G2 after a last pass over H; lane == row, so every global
// store of a warp is one full 128-byte line of the tau-major result.  The warp-per-row kernel
// (multitau_warp.cu) spends most of its instructions on moving data between lanes (scans, ballots,
// reductions, staging for coalesced output); here there is none of that, and quantities that depend
// on the delay only (thresholds, scales, divisors) are computed once per warp for 32 rows.
// In compat mode three short cooperative phases come first: merge-level histogram,
// live counts + first stale slot per level, the threshold key K* where the cheap filter fires.
#include <algorithm>
#include <cstdlib>

#include "internal.h"

#include "multitau_slice_coop.h"

namespace xpcs {

using sl::SlSched;

// -DXPCS_SL_TRACE: a few CTAs print where their warps spend their cycles (diagnostics; profiles/trace_slice.py)
#ifdef XPCS_SL_TRACE
#define SL_TRACE(what, id) do { if (lane == 0 && trn[warp] < 15) { trc[warp][trn[warp]] = ((long long)(what) << 56) | ((long long)((id) & 0xff) << 48) | (clock64() - t_start); trn[warp]++; } } while (0)
#else
#define SL_TRACE(what, id) do { } while (0)
#endif

constexpr int kSlMaxWarps = 16;
constexpr uint32_t kSlFull = 0xffffffffu;

struct SlArgs {
    unsigned char *fallback;   // [n_slices]
    int len_cap;               // longest slice handled here
    int np, nps, nd, nio;      // pieces of the pair walk (large, small), of the on-the-fly dense walk, of the IF and of the IP walk
    int ld_factor;             // dense levels start where L_l <= ld_factor * (longest row of the slice) ...
    int ld_cap;                // ... but not beyond this level (H is sized for the sparse slots up to it)
    int h_rows;                // rows of H: the delay slots of the levels below ld_cap
    int bins_rows;             // rows of the first 8-bit bin array (0: no bin arrays, dense levels on the fly)
    int area_words;            // the work area behind the tables (see sl_area_words)
    SlSched s;
};

constexpr int kSlHdr = 96;     // rlen[32], tot[32], smin[32]

// The work area holds, one after the other in time: the compat scratch tables (merge-level histogram and first
// stale slots, dead once the limits are known), then either the two 8-bit bin arrays or -- for slices that take
// the on-the-fly walk -- the G2 numerators of the dense levels.
static inline size_t sl_area_words(int T, int nl, int h_rows, bool compat, int bins_rows)
{
    size_t w = (size_t)std::min(T, T - h_rows + 8) * 32;  // a slice may start its dense levels one level below ld_cap
    if (compat) w = std::max(w, (size_t)(sl::kMlRows + nl) * 32);
    if (bins_rows > 0) w = std::max(w, (size_t)(bins_rows + (bins_rows >> 1) + (bins_rows >> 2)) * 8);
    return w;
}

// shared-memory words of a CTA
static inline size_t sl_smem_words(int len_cap, int T, int nl, int h_rows, bool compat, int bins_rows)
{
    return kSlHdr + (size_t)(len_cap + 1) * 32 + (size_t)h_rows * 32 + (size_t)nl * 32 + (compat ? (size_t)nl * 32 : 0) +
           sl_area_words(T, nl, h_rows, compat, bins_rows);
}

template <int DPL, bool COMPAT>
__global__ void __launch_bounds__(kSlMaxWarps * 32, 2) k_multitau_slice(MtArgs a, SlArgs m)
{
    extern __shared__ __align__(16) uint32_t sl_smem[];
    const int s = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int len = a.slice_len[s];
    if (len > m.len_cap) {  // CTA-uniform
        if (tid == 0) m.fallback[s] = 1;
        return;
    }
    const SlSched sc = m.s;
    const int T = sc.T, nl = sc.nl, F = sc.F;
#ifdef XPCS_SL_TRACE
    __shared__ long long trc[kSlMaxWarps][16];
    __shared__ int trn[kSlMaxWarps];
    const long long t_start = clock64();
    if (lane == 0) trn[warp] = 0;
    __syncwarp();
#endif
    uint32_t *rlen = sl_smem;                         // [32]
    uint32_t *tot = sl_smem + 32;                     // [32] sum of the counts of a row
    uint32_t *sminS = sl_smem + 64;                   // [32] compat: smallest first stale slot up to the first dense level
    uint32_t *evS = sl_smem + kSlHdr;                 // [len_cap + 1][32]
    uint32_t *H = evS + (size_t)(m.len_cap + 1) * 32; // [h_rows][32] G2 numerators of the sparse levels
    uint32_t *lim = H + (size_t)m.h_rows * 32;        // [nl][32] frame limit (l < ld) / key limit (l >= ld)
    uint32_t *nlive = lim + (size_t)nl * 32;          // compat: [nl][32] live bins per level
    uint32_t *area = nlive + (COMPAT ? (size_t)nl * 32 : 0);
    uint32_t *cntml = area;                           // compat, until the limits are known: [33][32]
    uint32_t *sbx = area + sl::kMlRows * 32;          //                                     [nl][32]
    // 8-bit bins, one array per dense task: [bins_rows][32], [bins_rows / 2][32], [bins_rows / 4][32] bytes
    const int boff1 = m.bins_rows * 8, boff2 = (m.bins_rows + (m.bins_rows >> 1)) * 8;
    uint32_t *Hd = area;                              // on-the-fly dense walk: numerators of the slots from hsp on
    __shared__ int qctr;

    // ---- the tile, as it lies in the store; words past the end of a row become sentinels
    if (tid < 32) {
        rlen[tid] = (uint32_t)a.row_len[s * kSlice + tid];
        tot[tid] = 0u;
    }
    if (tid == 0) qctr = 0;
    for (int t = tid; t < m.h_rows * 32; t += nthreads) H[t] = 0u;
    if (COMPAT)
        for (int t = tid; t < sl::kMlRows * 32; t += nthreads) cntml[t] = 0u;
    __syncthreads();
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(a.store) + a.slice_base[s]);
        uint4 *d = reinterpret_cast<uint4 *>(evS);
        const int q4 = (tid & 7) * 4;  // this thread always copies the rows q4 .. q4+3 (blockDim is a multiple of 8)
        const int n0 = (int)rlen[q4], n1 = (int)rlen[q4 + 1], n2 = (int)rlen[q4 + 2], n3 = (int)rlen[q4 + 3];
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        const int total4 = (len + 1) * 8;
        for (int idx = tid; idx < total4; idx += nthreads) {
            const int j = idx >> 3;
            uint4 v = make_uint4(sl::kSent, sl::kSent, sl::kSent, sl::kSent);
            if (j < len) v = g[idx];
            if (j < n0) c0 += v.x & sl::kCMask; else v.x = sl::kSent;
            if (j < n1) c1 += v.y & sl::kCMask; else v.y = sl::kSent;
            if (j < n2) c2 += v.z & sl::kCMask; else v.z = sl::kSent;
            if (j < n3) c3 += v.w & sl::kCMask; else v.w = sl::kSent;
            d[idx] = v;
        }
        atomicAdd(&tot[q4], c0);
        atomicAdd(&tot[q4 + 1], c1);
        atomicAdd(&tot[q4 + 2], c2);
        atomicAdd(&tot[q4 + 3], c3);
    }
    __syncthreads();
    SL_TRACE(1, 0);
    const uint32_t total = tot[lane];
    if (__any_sync(kSlFull, total >= 65536u)) {  // 32-bit numerators could overflow: the lane-per-row kernel redoes the slice
        if (tid == 0) m.fallback[s] = 1;         // (every warp sees the same 32 totals: the exit is CTA-uniform)
        return;
    }
    const int n = (int)rlen[lane];
    const uint32_t *ev = evS + lane;

    // first dense level: L_l <= ld_factor * (longest row), at most ld_cap.  CTA-uniform, so that the warps agree on
    // who does what
    int ld = min(nl, m.ld_cap);
    for (int l = 1; l < ld; l++)
        if ((F >> l) <= m.ld_factor * max(len, 1)) {
            ld = l;
            break;
        }
    const int hsp = min(T, sc.cnt0 + DPL * (ld - 1));  // delay slots of the sparse levels (<= h_rows)
    // 8-bit bin arrays for the dense levels when no row's counts sum beyond 255 (CTA-uniform), else the on-the-fly walk
    const bool use8 = m.bins_rows > 0 && ld <= sc.lastl && (F >> ld) <= m.bins_rows && __all_sync(kSlFull, total <= 255u);

    // ---- per-level limits
    if (COMPAT) {
        {
            const int chunk = (len + nwarps - 1) / nwarps;
            const int i0 = warp * chunk;
            sl::lane_mlhist(ev, i0, min(n, i0 + chunk), cntml + lane);
        }
        SL_TRACE(2, 0);
        __syncthreads();
        for (int l = 1 + warp; l <= sc.lastl; l += nwarps)
            sl::lane_level_base(ev, n, l, ld, F, cntml + lane, nlive + lane, sbx + lane, !use8);
        if (warp == 0) {
            nlive[lane] = (uint32_t)n;
            sbx[lane] = (uint32_t)sl::kInfKey;
        }
        SL_TRACE(3, 0);
        __syncthreads();
        for (int l = warp; l < nl; l += nwarps) {
            const int Ll = F >> l;
            uint32_t v = l < ld ? (uint32_t)Ll << l : (uint32_t)Ll;
            if (l >= 1 && l <= sc.lastl && (!use8 || l <= ld)) v = sl::lane_level_limit(ev, n, l, ld, F, nlive + lane, sbx + lane);
            lim[l * 32 + lane] = v;
        }
        if (warp == nwarps - 1) {
            int sm = sl::kInfKey;
            for (int j = 1; j <= min(ld, sc.lastl); j++) sm = min(sm, (int)sbx[j * 32 + lane]);
            sminS[lane] = (uint32_t)sm;
        }
    } else {
        for (int l = warp; l < nl; l += nwarps) {
            const int Ll = F >> l;
            lim[l * 32 + lane] = l < ld ? (uint32_t)Ll << l : (uint32_t)Ll;
        }
    }
    SL_TRACE(4, 0);
    __syncthreads();
    SL_TRACE(5, use8);
    if (!use8 && hsp < T) {  // CTA-uniform: the scratch tables are dead, the area now holds the dense numerators
        for (int t = tid; t < (T - hsp) * 32; t += nthreads) Hd[t] = 0u;
        __syncthreads();
    }

    // ---- the tasks, largest first, taken by whichever warp is free
    const int64_t r = (int64_t)s * kSlice + lane;
    const int ntd = ld > sc.lastl ? 0 : (use8 ? 3 : m.nd);
    const int ntasks = ntd + 2 * m.nio + m.np + m.nps;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&qctr, 1);
        t = __shfl_sync(kSlFull, t, 0);
        if (t >= ntasks) break;
        SL_TRACE(6, t);
        if (t < ntd) {
            if (use8) {
                if (ld + t > sc.lastl) continue;
                uint8_t *B = reinterpret_cast<uint8_t *>(area + (t == 0 ? 0 : (t == 1 ? boff1 : boff2)));
                const int rows = F >> (ld + t);
                for (int idx = lane; idx < rows * 8; idx += 32) reinterpret_cast<uint32_t *>(B)[idx] = 0u;
                __syncwarp();
                sl::lane_dense8<DPL, COMPAT>(t, ev, n, ld, sc, B + lane, lim + lane, nlive + lane,
                                             COMPAT ? (int)sminS[lane] : sl::kInfKey, a.G2 + r, a.R_pad);
            } else {
                constexpr int W = 2 * DPL + 1;
                int bins = 0;
                for (int l = ld; l <= sc.lastl; l++) bins += F >> l;
                int target = ((bins + m.nd - 1) / m.nd + W - 1) / W * W;
                if (target < W) target = W;
                int piece = 0;
                for (int l = ld; l <= sc.lastl; l++) {
                    const int Ll = F >> l;
                    const int cnt = sl::level_count<DPL>(sc, l);
                    const int npieces = max(1, (Ll + target - 1) / target);
                    for (int k = 0; k < npieces; k++, piece++) {
                        if (piece % m.nd != t) continue;
                        const int tb = k * target;
                        const int te = k == npieces - 1 ? Ll : tb + target;
                        uint32_t acc[DPL];
#pragma unroll
                        for (int d = 0; d < DPL; d++) acc[d] = 0u;
                        sl::lane_dense<DPL>(ev, n, l, tb, te, (int)lim[l * 32 + lane], acc);
#pragma unroll
                        for (int d = 0; d < DPL; d++)
                            if (d < cnt) atomicAdd(&Hd[(sc.cnt0 + (l - 1) * DPL + d - hsp) * 32 + lane], acc[d]);
                    }
                }
            }
        } else if (t < ntd + 2 * m.nio) {
            const int q = t - ntd;
            const int part = q >> 1;
            const int ta = (int)((int64_t)T * part / m.nio), tb = (int)((int64_t)T * (part + 1) / m.nio);
            if (q & 1) sl::lane_ip<DPL>(ev, n, total, sc, ta, tb, a.IP + (int64_t)ta * a.R_pad + r, a.R_pad);
            else sl::lane_if<DPL>(ev, n, total, sc, ta, tb, a.IF + (int64_t)ta * a.R_pad + r, a.R_pad);
        } else {
            // np pieces deal out the first three quarters of a row's events, nps small ones the rest: the last tasks
            // of the queue are short, so the warps finish close to each other
            int piece = t - ntd - 2 * m.nio;
            const int cut = m.nps > 0 ? n - (n >> 2) : n;
            int ia = 0, ib = cut, istep = m.np;
            if (piece >= m.np) {
                piece -= m.np;
                ia = cut;
                ib = n;
                istep = m.nps;
            }
            if (ld - 1 < sc.lastl) sl::lane_pairs<DPL, true>(ev, ia, ib, piece, istep, ld, sc, lim + lane, H + lane);
            else sl::lane_pairs<DPL, false>(ev, ia, ib, piece, istep, ld, sc, lim + lane, H + lane);
        }
    }
    SL_TRACE(7, 0);
    __syncthreads();
    SL_TRACE(8, 0);

    // ---- G2: one division per slot (the 8-bit bin tasks have written theirs)
    const int tend = use8 ? hsp : T;
    for (int ti = warp; ti < tend; ti += nwarps) {
        const uint32_t num = ti < hsp ? H[ti * 32 + lane] : Hd[(ti - hsp) * 32 + lane];
        a.G2[(int64_t)ti * a.R_pad + r] = sl::g2_value<DPL>(num, ti, sc);
    }
    SL_TRACE(9, 0);
#ifdef XPCS_SL_TRACE
    if ((s % 1031) == 7 && lane == 0)
        for (int k = 0; k < trn[warp]; k++)
            printf("SLT %d %d %d %d %lld\n", s, warp, (int)(trc[warp][k] >> 56), (int)((trc[warp][k] >> 48) & 0xff),
                   trc[warp][k] & 0xffffffffffffLL);
#endif
}

template <int DPL, bool COMPAT>
static int run_slice(xpcs_handle_s *h, MtArgs &a, SlArgs &m, size_t bytes, int warps)
{
    int rc = check_cuda(h, cudaFuncSetAttribute(k_multitau_slice<DPL, COMPAT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)bytes), "multitau_slice smem attr");
    if (rc) return rc;
    LaunchScope ls(h, "k_multitau_slice");
    k_multitau_slice<DPL, COMPAT><<<h->n_slices, warps * 32, bytes, h->stream>>>(a, m);
    return XPCS_OK;
}

static int sl_len_cap(const xpcs_handle_s *h)
{
    // a few outlier rows (hot pixels) must not dictate the shared-memory budget of every CTA: slices more than
    // four times longer than the mean slice go to the lane-per-row kernel
    int len_cap = h->max_row > 0 ? h->max_row : 1;
    const int64_t mean_len = h->n_slices > 0 ? h->store_words / kSlice / h->n_slices : 0;
    return (int)std::min<int64_t>(len_cap, std::max<int64_t>(256, 4 * mean_len));
}

static int sl_env(const char *name, int lo, int hi, int dflt)
{
    if (const char *e = getenv(name)) {
        const int q = atoi(e);
        if (q >= lo && q <= hi) return q;
    }
    return dflt;
}

// rows of the first 8-bit bin array: the most bins any slice can have at its first dense level
static int sl_bins_rows(const xpcs_handle_s *h, int len_cap, int ld_factor)
{
    if (!sl_env("XPCS_SL_BINS8", 0, 1, 1)) return 0;
    const Sched &sc = h->sched;
    for (int l = 1; l < sc.n_levels; l++)
        if ((sc.frames >> l) <= ld_factor * std::max(len_cap, 1)) return sc.count[l] > 0 ? (sc.frames >> l) : 0;
    return 0;
}

// first dense level of the longest slice this launch handles, and the cap of the kernel (one level beyond it)
static int sl_ld_min(const xpcs_handle_s *h, int len_cap, int ld_factor)
{
    const Sched &sc = h->sched;
    for (int l = 1; l < sc.n_levels; l++)
        if ((sc.frames >> l) <= ld_factor * std::max(len_cap, 1)) return l;
    return sc.n_levels;
}

static int sl_h_rows(const xpcs_handle_s *h, int ld_cap)
{
    return std::min(h->T, h->sched.count[0] + h->prm.delays_per_level * (ld_cap - 1));
}

static size_t sl_bytes(const xpcs_handle_s *h, int len_cap, bool compat, int ld_factor, int bins_rows)
{
    const int ld_cap = std::min(h->sched.n_levels, sl_ld_min(h, len_cap, ld_factor) + 1);
    return 4 * sl_smem_words(len_cap, h->T, h->sched.n_levels, sl_h_rows(h, ld_cap), compat, bins_rows);
}

// The slice kernel covers what the warp kernel covers (integer counts, dpl 4 or 8, frames below 2^20, the
// regular schedule) as long as three CTAs share an SM; longer rows stay with the warp-per-row kernel.
bool multitau_slice_eligible(const xpcs_handle_s *h)
{
    if (!multitau_warp_eligible(h)) return false;
    if (const char *e = getenv("XPCS_MT_KERNEL")) {  // diagnostics: "warp" or "slice"
        if (e[0] == 'w') return false;
        if (e[0] == 's') return true;
    }
    const bool compat = (h->prm.compat_flags & XPCS_COMPAT_STALE_TAIL) != 0;
    return sl_bytes(h, sl_len_cap(h), compat, 4, 0) <= 56 * 1024;
}

int launch_multitau_slice(xpcs_handle_s *h, MtArgs &a)
{
    int rc = ensure(h, h->d_mt_fallback, (size_t)(h->n_slices > 0 ? h->n_slices : 1), "multitau fallback flags");
    if (rc) return rc;
    cudaMemsetAsync(h->d_mt_fallback.p, 0, (size_t)(h->n_slices > 0 ? h->n_slices : 1), h->stream);
    if (h->n_slices == 0) return XPCS_OK;
    int smem_cap = 0;
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    const bool compat = a.compat != 0;
    SlArgs m{};
    m.fallback = h->d_mt_fallback.p;
    const Sched &sc = h->sched;
    m.s.F = sc.frames;
    m.s.nl = sc.n_levels;
    m.s.T = h->T;
    m.s.cnt0 = sc.count[0];
    for (int l = 1; l < sc.n_levels; l++)
        if (sc.count[l] > 0) {
            m.s.lastl = l;
            m.s.cnt_last = sc.count[l];
        }
    m.ld_factor = sl_env("XPCS_SL_LD", 1, 16, 4);
    int len_cap = sl_len_cap(h);
    while (len_cap > 1 && sl_bytes(h, len_cap, compat, m.ld_factor, 0) > (size_t)smem_cap) len_cap = len_cap * 3 / 4;
    if (sl_bytes(h, len_cap, compat, m.ld_factor, 0) > (size_t)smem_cap) {  // T too large: everything falls back
        cudaMemsetAsync(h->d_mt_fallback.p, 1, (size_t)h->n_slices, h->stream);
        return XPCS_OK;
    }
    m.len_cap = len_cap;
    m.ld_cap = std::min(sc.n_levels, sl_ld_min(h, len_cap, m.ld_factor) + 1);
    m.h_rows = sl_h_rows(h, m.ld_cap);
    // 8-bit bin arrays for the dense levels as long as three CTAs still share an SM
    m.bins_rows = sl_bins_rows(h, len_cap, m.ld_factor);
    if (sl_bytes(h, len_cap, compat, m.ld_factor, m.bins_rows) > (size_t)(smem_cap + 1024) / 3 - 1024) m.bins_rows = 0;
    m.area_words = (int)sl_area_words(h->T, sc.n_levels, m.h_rows, compat, m.bins_rows);
    m.np = sl_env("XPCS_SL_PAIR_PIECES", 1, 32, 8);    // diagnostics
    m.nps = sl_env("XPCS_SL_PAIR_TAIL", 0, 32, 8);
    m.nd = sl_env("XPCS_SL_DENSE_PIECES", 1, 16, 4);
    m.nio = sl_env("XPCS_SL_IO_PIECES", 1, 8, 2);
    const int warps = sl_env("XPCS_SL_WARPS", 2, kSlMaxWarps, 8);
    const size_t bytes = sl_bytes(h, len_cap, compat, m.ld_factor, m.bins_rows);
    const int dpl = h->prm.delays_per_level;
    if (dpl == 8) rc = compat ? run_slice<8, true>(h, a, m, bytes, warps) : run_slice<8, false>(h, a, m, bytes, warps);
    else rc = compat ? run_slice<4, true>(h, a, m, bytes, warps) : run_slice<4, false>(h, a, m, bytes, warps);
    if (rc) return rc;
    return check_cuda(h, cudaGetLastError(), "k_multitau_slice");
}

}  // namespace xpcs
