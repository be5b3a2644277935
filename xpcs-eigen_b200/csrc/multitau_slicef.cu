// multitau_slicef.cu -- multi-tau correlator for float-valued rows: lane = pixel row, warps = tasks.
//
// Replaces Corr::multiTau2 (reference corr.cpp:315-431) for the float store (word = frame << 32 | float bits:
// flat-fielded, averaged, frame-sum-normalised or dense-source data) when the 32 rows of a slice fit a
// shared-memory tile (rows of up to a few hundred events: BASELINE configs[1] after the threshold, ~3 * 10^2 per row).
// Same mapping as k_multitau_slice (multitau_slice.cu): one CTA owns one slice, the tile is copied into shared
// memory once (16-byte loads, frames and values into two planes; bank == lane afterwards) and the warps then pull
// DIFFERENT jobs on the same 32 rows from a queue, every lane on its own row with plain sequential code
// (multitau_slicef_core.h -- the same source is compiled for the host and checked against the oracle on the CPU):
//     pair pieces    G2 of the sparse levels: one flattened walk over the event pairs of a row, events dealt
//                    round-robin to the pieces; every piece adds into its OWN accumulator array (there is no native
//                    shared-memory float add, and piece-private sums make the result independent of which warp
//                    ran which piece)
//     dense pieces   G2 of the dense levels, (level, bin range) pieces: the bins are formed on the fly under a
//                    sliding register window, fp64; every piece leaves its sums in its own slot
//     IF / IP parts  forward / backward walk with an fp64 running sum at the ascending / descending thresholds
// IP and IF leave straight from registers, G2 after a last pass that adds the pieces up in a fixed order; lane == row,
// so every global store of a warp is one full 128-byte line of the tau-major result.  The warp-per-row kernel it
// replaces for these rows (multitau_warpf.cu) spends its time in fp64 shared-memory CAS loops, scans and shuffles.
// Slices longer than the tile (hot pixels) are flagged; k_multitau_warpf and then the lane-per-row kernel take them.
// In compat mode the same three cooperative phases as in k_multitau_slice come first.
#include <algorithm>
#include <cstdlib>

#include "internal.h"

#define XS_NS slf
#define XS_CB 0
#include "multitau_slice_coop.h"
#include "multitau_slicef_core.h"

namespace xpcs {

using slf::SlSched;

// -DXPCS_SL_TRACE: a few CTAs print where their warps spend their cycles (diagnostics; profiles/trace_slice.py)
#ifdef XPCS_SL_TRACE
#define SF_TRACE(what, id) do { if (lane == 0 && trn[warp] < 31) { trc[warp][trn[warp]] = ((long long)(what) << 56) | ((long long)((id) & 0xff) << 48) | (clock64() - t_start); trn[warp]++; } } while (0)
#else
#define SF_TRACE(what, id) do { } while (0)
#endif

constexpr int kSfMaxWarps = 24;
constexpr int kSfMaxPairPieces = 16;
constexpr uint32_t kSfFull = 0xffffffffu;
constexpr int kSfHdr = 128;  // rlen[32], tot[32] (double), pcs[32]

struct SfArgs {
    unsigned char *fallback;   // [n_slices]
    int len_cap;               // longest slice handled here
    int np, nps, nd, nio;      // pair pieces (large, small), target number of dense pieces, parts of the IF and of the IP walk
    int ld_factor, ld_cap;     // dense levels start where L_l <= ld_factor * (longest row of the slice), at most at ld_cap
    int h_rows;                // rows of a pair piece's accumulator array: the delay slots of the levels below ld_cap
    int dp_rows;               // dense piece slots available
    int io_first;              // queue order: pair pieces, IF / IP parts, dense pieces (else the IF / IP parts last)
    SlSched s;
};

static inline size_t sf_area_words(int nl, int dp_rows, int dpl, bool compat)
{
    size_t w = (size_t)dp_rows * dpl * 32;
    if (compat) w = std::max(w, (size_t)(slf::kMlRows + nl) * 32);
    return w;
}

static inline size_t sf_smem_words(int len_cap, int nl, int h_rows, int pieces, int dp_rows, int dpl, bool compat)
{
    return kSfHdr + (size_t)2 * (len_cap + 1) * 32 + (size_t)pieces * h_rows * 32 + (size_t)nl * 32 +
           (compat ? (size_t)nl * 32 : 0) + sf_area_words(nl, dp_rows, dpl, compat);
}

template <int DPL, bool COMPAT>
__global__ void __launch_bounds__(kSfMaxWarps * 32, 1) k_multitau_slicef(MtArgs a, SfArgs m)
{
    extern __shared__ __align__(16) uint32_t sf_smem[];
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
    const int npieces = m.np + m.nps;
    uint32_t *rlen = sf_smem;                                         // [32]
    double *tot = reinterpret_cast<double *>(sf_smem + 32);           // [32] sum of the values of a row
    int *pcs = reinterpret_cast<int *>(sf_smem + 96);                 // [32] dense pieces of level l (0 outside ld .. lastl)
    uint32_t *frS = sf_smem + kSfHdr;                                 // [len_cap + 1][32] frames
    float *vlS = reinterpret_cast<float *>(frS + (size_t)(m.len_cap + 1) * 32);  // [len_cap + 1][32] values
    float *Hp = vlS + (size_t)(m.len_cap + 1) * 32;                   // [npieces][h_rows][32] pair sums of the sparse levels
    uint32_t *lim = reinterpret_cast<uint32_t *>(Hp + (size_t)npieces * m.h_rows * 32);  // [nl][32] frame limit (l < ld) / key limit
    uint32_t *nlive = lim + (size_t)nl * 32;                          // compat: [nl][32] live bins per level
    uint32_t *area = nlive + (COMPAT ? (size_t)nl * 32 : 0);
    uint32_t *cntml = area;                                           // compat, until the limits are known: [33][32]
    uint32_t *sbx = area + slf::kMlRows * 32;                         //                                     [nl][32]
    float *Dp = reinterpret_cast<float *>(area);                      // afterwards: [dp_rows][DPL][32] sums of the dense pieces
    __shared__ int qctr;
#ifdef XPCS_SL_TRACE
    __shared__ long long trc[kSfMaxWarps][32];
    __shared__ int trn[kSfMaxWarps];
    const long long t_start = clock64();
    if (lane == 0) trn[warp] = 0;
    __syncwarp();
#endif

    // ---- the tile: a word of the store is value | frame << 32, so a 16-byte load carries rows 2q and 2q+1 of one step
    if (tid < 32) rlen[tid] = (uint32_t)a.row_len[s * kSlice + tid];
    if (tid == 0) qctr = 0;
    for (int t = tid; t < npieces * m.h_rows * 32; t += nthreads) Hp[t] = 0.0f;
    if (COMPAT)
        for (int t = tid; t < slf::kMlRows * 32; t += nthreads) cntml[t] = 0u;
    __syncthreads();
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned long long *>(a.store) + a.slice_base[s]);
        const int q2 = (tid & 15) * 2;  // this thread always copies the rows q2, q2+1 (blockDim is a multiple of 16)
        const int n0 = (int)rlen[q2], n1 = (int)rlen[q2 + 1];
        const int total16 = (len + 1) * 16;
        for (int idx = tid; idx < total16; idx += nthreads) {
            const int j = idx >> 4;
            uint4 v = make_uint4(0u, slf::kSent, 0u, slf::kSent);
            if (j < len) v = g[idx];
            const uint2 f = make_uint2(j < n0 ? v.y : slf::kSent, j < n1 ? v.w : slf::kSent);
            *reinterpret_cast<uint2 *>(frS + 2 * idx) = f;
            *reinterpret_cast<uint2 *>(vlS + 2 * idx) = make_uint2(v.x, v.z);
        }
    }
    __syncthreads();
    SF_TRACE(1, 0);
    const int n = (int)rlen[lane];
    const uint32_t *fr = frS + lane;
    const float *vl = vlS + lane;

    // first dense level: L_l <= ld_factor * (longest row), at most ld_cap.  CTA-uniform
    int ld = min(nl, m.ld_cap);
    for (int l = 1; l < ld; l++)
        if ((F >> l) <= m.ld_factor * max(len, 1)) {
            ld = l;
            break;
        }
    const int hsp = min(T, sc.cnt0 + DPL * (ld - 1));  // delay slots of the sparse levels (<= h_rows)
    const int target = ld <= sc.lastl ? slf::dense_target<DPL>(sc, ld, m.nd) : 0;
    if (tid < 32) pcs[tid] = (tid >= ld && tid <= sc.lastl) ? slf::dense_pieces(sc, tid, target) : 0;

    // ---- row sums (last warp) and per-level limits
    if (warp == nwarps - 1) tot[lane] = slf::lanef_total(vl, n);
    if (COMPAT) {
        {
            const int chunk = (len + nwarps - 1) / nwarps;
            const int i0 = warp * chunk;
            slf::lane_mlhist(fr, i0, min(n, i0 + chunk), cntml + lane);
        }
        SF_TRACE(2, 0);
        __syncthreads();
        for (int l = 1 + warp; l <= sc.lastl; l += nwarps)
            slf::lane_level_base(fr, n, l, ld, F, cntml + lane, nlive + lane, sbx + lane, true);
        if (warp == 0) {
            nlive[lane] = (uint32_t)n;
            sbx[lane] = (uint32_t)slf::kInfKey;
        }
        SF_TRACE(3, 0);
        __syncthreads();
        for (int l = warp; l < nl; l += nwarps) {
            const int Ll = F >> l;
            uint32_t v = l < ld ? (uint32_t)Ll << l : (uint32_t)Ll;
            if (l >= 1 && l <= sc.lastl) v = slf::lane_level_limit(fr, n, l, ld, F, nlive + lane, sbx + lane);
            lim[l * 32 + lane] = v;
        }
    } else {
        for (int l = warp; l < nl; l += nwarps) {
            const int Ll = F >> l;
            lim[l * 32 + lane] = l < ld ? (uint32_t)Ll << l : (uint32_t)Ll;
        }
    }
    SF_TRACE(4, 0);
    __syncthreads();  // (the scratch tables are dead from here on: the area holds the sums of the dense pieces)
    SF_TRACE(5, 0);
    const double total = tot[lane];

    // ---- the tasks, taken by whichever warp is free: pair pieces, dense pieces, IF / IP parts
    const int64_t r = (int64_t)s * kSlice + lane;
    int ndp = 0;
    for (int l = ld; l <= sc.lastl; l++) ndp += pcs[l];
    const int ntasks = ndp + npieces + 2 * m.nio;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&qctr, 1);
        t = __shfl_sync(kSfFull, t, 0);
        if (t >= ntasks) break;
        SF_TRACE(6, t);
        // queue order: the pair pieces (the longest tasks: their number is bounded by the accumulator arrays), the IF / IP
        // parts (the ones of the deep levels walk the whole row), then the dense pieces, whose last ones -- the deep levels
        // -- are the shortest tasks, so that the warps finish close to each other
        const int nio2 = 2 * m.nio;
        int kind, idx;  // 0 pair piece, 1 dense piece, 2 IF / IP part
        if (t < npieces) {
            kind = 0;
            idx = t;
        } else if (m.io_first) {
            kind = t < npieces + nio2 ? 2 : 1;
            idx = t - npieces - (kind == 1 ? nio2 : 0);
        } else {
            kind = t < npieces + ndp ? 1 : 2;
            idx = t - npieces - (kind == 2 ? ndp : 0);
        }
        if (kind == 0) {
            // np pieces deal out the first three quarters of a row's events, nps small ones the rest
            const int w = idx;
            int piece = w;
            const int cut = m.nps > 0 ? n - (n >> 2) : n;
            int ia = 0, ib = cut, istep = m.np;
            if (piece >= m.np) {
                piece -= m.np;
                ia = cut;
                ib = n;
                istep = m.nps;
            }
            float *H = Hp + (size_t)w * m.h_rows * 32 + lane;
            if (ld - 1 < sc.lastl) slf::lanef_pairs<DPL, true>(fr, vl, ia, ib, piece, istep, ld, sc, lim + lane, H);
            else slf::lanef_pairs<DPL, false>(fr, vl, ia, ib, piece, istep, ld, sc, lim + lane, H);
        } else if (kind == 1) {
            const int p = idx;
            int l = ld, k = p;
            for (; k >= pcs[l]; l++) k -= pcs[l];
            const int Ll = F >> l;
            const int tb = k * target;
            const int te = k == pcs[l] - 1 ? Ll : tb + target;
            double acc[DPL];
#pragma unroll
            for (int d = 0; d < DPL; d++) acc[d] = 0.0;
            slf::lanef_dense<DPL>(fr, vl, n, l, tb, te, (int)lim[l * 32 + lane], acc);
#pragma unroll
            for (int d = 0; d < DPL; d++) Dp[(p * DPL + d) * 32 + lane] = (float)acc[d];
        } else {
            // the parts of the deep levels first: they are the long ones
            const int q = nio2 - 1 - idx;
            const int part = q >> 1;
            const int ta = (int)((int64_t)T * part / m.nio), tb = (int)((int64_t)T * (part + 1) / m.nio);
            if (q & 1) slf::lanef_ip<DPL>(fr, vl, n, total, sc, ta, tb, a.IP + (int64_t)ta * a.R_pad + r, a.R_pad);
            else slf::lanef_if<DPL>(fr, vl, n, total, sc, ta, tb, a.IF + (int64_t)ta * a.R_pad + r, a.R_pad);
        }
    }
    SF_TRACE(7, 0);
    __syncthreads();
    SF_TRACE(8, 0);

    // ---- G2: the pieces added up in a fixed order (fp64), one rounding to fp32 and one division per slot
    for (int ti = warp; ti < T; ti += nwarps) {
        double num = 0.0;
        if (ti < hsp) {
            for (int w = 0; w < npieces; w++) num += (double)Hp[((size_t)w * m.h_rows + ti) * 32 + lane];
        } else {
            const int q = ti - sc.cnt0, l = 1 + q / DPL, d = q % DPL;
            int base = 0;
            for (int j = ld; j < l; j++) base += pcs[j];
            const int np_l = pcs[l];
            for (int k = 0; k < np_l; k++) num += (double)Dp[((base + k) * DPL + d) * 32 + lane];
        }
        a.G2[(int64_t)ti * a.R_pad + r] = slf::g2f_value<DPL>(num, ti, sc);
    }
    SF_TRACE(9, 0);
#ifdef XPCS_SL_TRACE
    if ((s % 1031) == 7 && lane == 0)
        for (int k = 0; k < trn[warp]; k++)
            printf("SLT %d %d %d %d %lld\n", s, warp, (int)(trc[warp][k] >> 56), (int)((trc[warp][k] >> 48) & 0xff),
                   trc[warp][k] & 0xffffffffffffLL);
#endif
}

template <int DPL, bool COMPAT>
static int run_slicef(xpcs_handle_s *h, MtArgs &a, SfArgs &m, size_t bytes, int warps)
{
    int rc = check_cuda(h, cudaFuncSetAttribute(k_multitau_slicef<DPL, COMPAT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)bytes), "multitau_slicef smem attr");
    if (rc) return rc;
    LaunchScope ls(h, "k_multitau_slicef");
    k_multitau_slicef<DPL, COMPAT><<<h->n_slices, warps * 32, bytes, h->stream>>>(a, m);
    return XPCS_OK;
}

static int sf_env(const char *name, int lo, int hi, int dflt)
{
    if (const char *e = getenv(name)) {
        const int q = atoi(e);
        if (q >= lo && q <= hi) return q;
    }
    return dflt;
}

// what a launch would look like; false if the kernel cannot take the typical slice of this job
static bool sf_plan(const xpcs_handle_s *h, SfArgs &m, size_t &bytes, int &warps)
{
    const Sched &sc = h->sched;
    const int dpl = h->prm.delays_per_level;
    const bool compat = (h->prm.compat_flags & XPCS_COMPAT_STALE_TAIL) != 0;
    int smem_cap = 0;
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    smem_cap -= 1024;  // (static shared memory of the kernel, reserve)
    m = SfArgs{};
    m.s.F = sc.frames;
    m.s.nl = sc.n_levels;
    m.s.T = h->T;
    m.s.cnt0 = sc.count[0];
    for (int l = 1; l < sc.n_levels; l++)
        if (sc.count[l] > 0) {
            m.s.lastl = l;
            m.s.cnt_last = sc.count[l];
        }
    // measured on C2 (profiles/r03_sweep_slicef.txt): the pair pieces are the longest tasks and their number is
    // bounded by the accumulator arrays, so the dense levels start one level earlier than in k_multitau_slice
    // (L_l <= 8 len: half the pairs), in a dozen pieces
    m.ld_factor = sf_env("XPCS_SF_LD", 1, 16, 8);
    m.nd = sf_env("XPCS_SF_DENSE_PIECES", 1, 64, 12);
    m.nio = sf_env("XPCS_SF_IO_PIECES", 1, 8, 2);
    m.np = sf_env("XPCS_SF_PAIR_PIECES", 1, kSfMaxPairPieces, 6);
    m.nps = sf_env("XPCS_SF_PAIR_TAIL", 0, kSfMaxPairPieces - m.np, 2);
    m.dp_rows = m.nd + sc.n_levels;
    m.io_first = sf_env("XPCS_SF_IO_FIRST", 0, 1, 0);  // (measured on C2: 20.17 ms with, 20.00 ms without)
    warps = sf_env("XPCS_SF_WARPS", 2, kSfMaxWarps, 24);
    // The rows beyond ~1000 events stay with the lane-per-row kernel, which keeps the reference's sequential fp32
    // order (multitau_warpf.cu: kMfExactLen); a few outlier rows (hot pixels) must not dictate the tile of every
    // CTA: slices more than four times longer than the mean slice are left to the kernels behind this one
    int len_cap = h->max_row > 0 ? std::min(h->max_row, 1024) : 1;
    const int64_t mean_len = h->n_slices > 0 ? h->store_words / kSlice / h->n_slices : 0;
    len_cap = (int)std::min<int64_t>(len_cap, std::max<int64_t>(64, 4 * mean_len));
    auto plan_bytes = [&](int lc) {
        int ld_min = sc.n_levels;
        for (int l = 1; l < sc.n_levels; l++)
            if ((sc.frames >> l) <= m.ld_factor * std::max(lc, 1)) {
                ld_min = l;
                break;
            }
        m.ld_cap = std::min(sc.n_levels, ld_min + 1);
        m.h_rows = std::min(h->T, sc.count[0] + dpl * (m.ld_cap - 1));
        return 4 * sf_smem_words(lc, sc.n_levels, m.h_rows, m.np + m.nps, m.dp_rows, dpl, compat);
    };
    // the typical slice (twice the mean) must fit, else this is not the kernel for the job
    const int typical = (int)std::min<int64_t>(len_cap, std::max<int64_t>(1, 2 * mean_len));
    if (plan_bytes(typical) > (size_t)smem_cap) return false;
    while (len_cap > typical && plan_bytes(len_cap) > (size_t)smem_cap) len_cap = std::max(typical, len_cap * 7 / 8);
    bytes = plan_bytes(len_cap);
    m.len_cap = len_cap;
    return true;
}

// Float rows, dpl 4 or 8, the regular schedule (what the float warp-per-row kernel covers), frames that fit the
// walks' 32-bit arithmetic, and a tile that fits shared memory.
bool multitau_slicef_eligible(const xpcs_handle_s *h)
{
    if (!multitau_warpf_eligible(h)) return false;
    if (h->prm.frames >= (1 << 27)) return false;
    if (const char *e = getenv("XPCS_MTF_KERNEL")) {  // diagnostics: "warp" or "slice"
        if (e[0] == 'w') return false;
    }
    SfArgs m;
    size_t bytes;
    int warps;
    return sf_plan(h, m, bytes, warps);
}

int launch_multitau_slicef(xpcs_handle_s *h, MtArgs &a)
{
    int rc = ensure(h, h->d_mt_fallback, (size_t)(h->n_slices > 0 ? h->n_slices : 1), "multitau fallback flags");
    if (rc) return rc;
    cudaMemsetAsync(h->d_mt_fallback.p, 0, (size_t)(h->n_slices > 0 ? h->n_slices : 1), h->stream);
    if (h->n_slices == 0) return XPCS_OK;
    SfArgs m;
    size_t bytes;
    int warps;
    if (!sf_plan(h, m, bytes, warps)) {
        cudaMemsetAsync(h->d_mt_fallback.p, 1, (size_t)h->n_slices, h->stream);
        return XPCS_OK;
    }
    m.fallback = h->d_mt_fallback.p;
    const bool compat = a.compat != 0;
    const int dpl = h->prm.delays_per_level;
    if (dpl == 8) rc = compat ? run_slicef<8, true>(h, a, m, bytes, warps) : run_slicef<8, false>(h, a, m, bytes, warps);
    else rc = compat ? run_slicef<4, true>(h, a, m, bytes, warps) : run_slicef<4, false>(h, a, m, bytes, warps);
    if (rc) return rc;
    return check_cuda(h, cudaGetLastError(), "k_multitau_slicef");
}

}  // namespace xpcs
