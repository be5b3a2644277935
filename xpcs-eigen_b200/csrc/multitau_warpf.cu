// multitau_warpf.cu -- multi-tau correlator, one WARP per pixel row, float-valued rows.
//
// Replaces Corr::multiTau2 (reference corr.cpp:315-431) for the float store
// (word = frame << 32 | float bits): flat-fielded, averaged, frame-sum-normalised or
// dense-source (dark-subtracted, thresholded) data.  Same formulation as multitau_warp.cu:
// nothing is compacted level by level, every quantity is a function of the row's level-0
// events (f_i, v_i):
//   L_l = F >> l, lim_l = L_l << l;  an event is live at level l iff f < lim_l   (corr.cpp:349-390)
//   IP(l,t') = 2^-l PS((L_l - t') << l),  IF(l,t') = 2^-l (PS(lim_l) - PS(t' << l))  (corr.cpp:403,414-416)
//       PS(x) = sum of the values with f < x: fp64 prefix sums + binary search
//   G2, sparse levels (l < ld): one pass over event pairs i < j, products v_i v_j (exact in
//       fp64) accumulated in fp64 shared-memory accumulators;
//   G2, dense levels (l >= ld, first level with L_l <= 4 n): fp32 bin arrays
//       B_{l+1}[t] = B_l[2t] + B_l[2t+1] -- the reference's own cascade without its exact /2 --
//       and windowed products, lanes over t, fixed-shape reductions.
// One rounding to fp32 and one IEEE division per output.  The reference accumulates the same
// sums sequentially in fp32; for rows of up to ~1000 events the two agree to ~1e-6 relative
// (north_star tolerance 1e-5, tests/test_gpu_dense.py, tests/test_gpu_parity.py); longer rows
// are left to the lane-per-row kernel, which reproduces the reference's summation order.
// XPCS_COMPAT_STALE_TAIL depends on the frames only and is handled exactly as
// in multitau_warp.cu: live counts from the merge levels, the first stale slot of every level
// as a filter, the boundary walk with rank/select over the events when it can matter.
//
// A CTA owns one slice of 32 rows; each warp pulls its rows straight from the slice (the
// rows in flight at any moment are neighbours, so their 32-byte sectors are shared through
// L1), and the 32 x T x 3 results leave through a shared-memory stage as 128-byte lines.
#include <algorithm>

#include "internal.h"

namespace xpcs {

constexpr uint32_t kFullF = 0xffffffffu;
constexpr int kInfF = 0x7fffffff;
constexpr int kMfMaxWarps = 16;
constexpr int kMfExactLen = 1024;

struct MfArgs {
    unsigned char *fallback;   // [n_slices]
    int len_cap;               // longest row handled here
    int pitch_t;               // odd, >= T
    int warp_words;            // per-warp shared words
    int T;
    int bin_cap;               // bins a warp can hold (dense levels start at L_l <= min(4 n, bin_cap))
    int inplace;               // the flags are those of k_multitau_slicef: take the flagged slices only, clear what is done
};

__device__ __forceinline__ float mf_pow2_neg(int e) { return __int_as_float((127 - e) << 23); }
__device__ __forceinline__ double mf_pow2_neg_d(int e) { return __longlong_as_double((long long)(1023 - e) << 52); }
__device__ __forceinline__ float mf_scaled_div(float num, int neff)
{
    return neff > 0 ? __fdiv_rn(num, (float)neff) : num;
}

// first index in [0, n) whose frame is >= f
__device__ __forceinline__ int mf_lower_bound(const uint32_t *fr, int n, uint32_t f)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (fr[mid] < f) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// event index of the p-th (0-based) event that starts a bin at `level`; warp-uniform
__device__ __forceinline__ int mf_select_head(const uint32_t *fr, int n, int level, int p, int lane)
{
    int base = 0;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int i = c0 + lane;
        bool head = false;
        if (i < n) head = (i == 0) || (((fr[i] ^ fr[i - 1]) >> level) != 0u);
        const unsigned mk = __ballot_sync(kFullF, head);
        const int c = __popc(mk);
        if (p < base + c) return c0 + (int)__fns(mk, 0, p - base + 1);
        base += c;
    }
    return n - 1;
}

// value the reference sees in slot p of its vector at `level`; warp-uniform
__device__ __forceinline__ int mf_key_at(const uint32_t *fr, int n, const uint32_t *nlive, int level, int p, int lane)
{
    int lv = level;
    if (p >= (int)nlive[level]) {
        lv = level - 1;
        while (lv > 0 && (int)nlive[lv] <= p) lv--;
    }
    const int i = mf_select_head(fr, n, lv, p, lane);
    return (int)(fr[i] >> lv);
}

// threshold key K*: targets with key >= K* are never found by the reference's search
__device__ __forceinline__ int mf_stale_threshold(const uint32_t *fr, int n, const uint32_t *nlive, int level, int lane)
{
    const int nl = (int)nlive[level];
    int first = 0, len = n;
    int curmin = kInfF;
    while (len > 0) {
        const int half = len >> 1;
        const int mid = first + half;
        if (mid >= nl) {
            curmin = min(curmin, mf_key_at(fr, n, nlive, level, mid, lane));
            len = half;
        } else {
            if (curmin != kInfF && mf_key_at(fr, n, nlive, level, mid, lane) > curmin) {
                const int k1 = mf_key_at(fr, n, nlive, level, first, lane);
                const int j = mf_lower_bound(fr, n, (uint32_t)(curmin + 1) << level);
                const int k2 = j < n ? (int)(fr[j] >> level) : kInfF;
                return max(k1, k2);
            }
            first = mid + 1;
            len = len - half - 1;
        }
    }
    return kInfF;
}

__device__ __forceinline__ double mf_shfl_up(double x, int o)
{
    return __shfl_up_sync(kFullF, x, o);
}

template <int DPL, bool COMPAT>
__global__ void __launch_bounds__(kMfMaxWarps * 32) k_multitau_warpf(MtArgs a, MfArgs m)
{
    constexpr int LG = DPL == 8 ? 3 : 2;
    constexpr int LO = DPL + 1;
    extern __shared__ __align__(16) uint32_t mf_smem[];
    const int s = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    if (m.inplace && !m.fallback[s]) return;  // done by k_multitau_slicef
    const int len = a.slice_len[s];
    if (len > m.len_cap) {  // CTA-uniform
        if (tid == 0) m.fallback[s] = 1;
        return;
    }
    uint32_t *outS = mf_smem;                                    // [3][32][pitch_t]
    uint32_t *wsm = outS + 3 * 32 * m.pitch_t + warp * m.warp_words;
    double *Hacc = reinterpret_cast<double *>(wsm);              // [T] G2 numerators
    double *totD = Hacc + m.T;                                   // [32]
    uint32_t *tb = reinterpret_cast<uint32_t *>(totD + 32);
    uint32_t *flim = tb, *cnts = tb + 32, *cntml = tb + 64, *nlive = tb + 96, *sminS = tb + 128;
    uint32_t *fr = tb + 160;                                     // [len_cap]
    float *vl = reinterpret_cast<float *>(fr + m.len_cap);       // [len_cap]
    uint32_t *U = fr + 2 * m.len_cap;                            // 8-byte aligned: ps (double), later bins (float)
    double *ps = reinterpret_cast<double *>(U);                  // [len_cap + 1]
    float *bh = reinterpret_cast<float *>(U);                    // [bin_cap + 64]

    const int F = a.sched.frames;
    const int nl = a.sched.n_levels;
    const int T = m.T;
    const int cnt0 = a.sched.count[0];
    const int lo0 = a.sched.lo[0];
    const unsigned long long *slice =
        reinterpret_cast<const unsigned long long *>(a.store) + a.slice_base[s];

    for (int rr = warp; rr < kSlice; rr += nwarps) {
        const int n = a.row_len[s * kSlice + rr];
        uint32_t *H = outS + rr * m.pitch_t;              // G2 floats
        uint32_t *oIP = H + 32 * m.pitch_t;
        uint32_t *oIF = oIP + 32 * m.pitch_t;

        // ---- phase 0: the row, frames and values apart
        for (int j = lane; j < n; j += 32) {
            const unsigned long long w = slice[(int64_t)j * kSlice + rr];
            fr[j] = (uint32_t)(w >> 32);
            vl[j] = __uint_as_float((uint32_t)w);
        }
        for (int t = lane; t < T; t += 32) Hacc[t] = 0.0;
        if (COMPAT) cntml[lane] = 0u;
        if (lane == 0) ps[0] = 0.0;
        __syncwarp();

        // ---- phase 1: fp64 prefix sums of the values, merge-level histogram
        double carry = 0.0;
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int i = c0 + lane;
            double x = i < n ? (double)vl[i] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = mf_shfl_up(x, o);
                if (lane >= o) x += y;
            }
            if (i < n) ps[i + 1] = carry + x;
            carry += __shfl_sync(kFullF, x, 31);
            if (COMPAT && i >= 1 && i < n) {
                const int ml = 32 - __clz((int)(fr[i] ^ fr[i - 1]));
                atomicAdd(&cntml[ml], 1u);
            }
        }
        __syncwarp();

        // first dense level: L_l <= 4 n (and the bins must fit the warp's area)
        int ld;
        {
            const unsigned mk =
                __ballot_sync(kFullF, lane >= 1 && lane < nl && (F >> lane) <= min(4 * max(n, 1), m.bin_cap));
            ld = mk ? (__ffs(mk) - 1) : nl;
        }

        // ---- phase 2: per-level tables (lane = level)
        {
            const int l = lane;
            const bool lv_ok = l < nl;
            const int Ll = lv_ok ? (F >> l) : 0;
            const int liml = lv_ok ? (Ll << l) : 0;
            const int cnt_l = lv_ok ? a.sched.count[l] : 0;
            totD[l] = ps[mf_lower_bound(fr, n, (uint32_t)liml)];
            cnts[l] = (l < ld) ? (uint32_t)cnt_l : 0u;
            uint32_t fl = (l < ld) ? (uint32_t)liml : 0u;
            if (COMPAT) {
                uint32_t ab = cntml[l];  // events that are not the first of their bin from level ml on
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(kFullF, ab, o);
                    if (lane >= o) ab += y;
                }
                const int dropped = (n > 0 && lv_ok && (int)(fr[n - 1] >> l) >= Ll) ? 1 : 0;
                const int nv = n - (int)ab - dropped;
                int prev = __shfl_up_sync(kFullF, nv, 1);
                if (lane == 0) prev = n;
                int sb = kInfF;
                if (l >= 1 && lv_ok && nv < prev) sb = (int)(fr[nv] >> (l - 1));
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(kFullF, sb, o);
                    if (lane >= o) sb = min(sb, y);
                }
                nlive[l] = (uint32_t)nv;
                sminS[l] = (uint32_t)sb;
                __syncwarp();
                const bool fire = l >= 1 && l < ld && cnt_l > 0 && nv < n && sb < Ll;
                unsigned mk = __ballot_sync(kFullF, fire);
                while (mk) {
                    const int lv = __ffs(mk) - 1;
                    mk &= mk - 1;
                    const int ks = mf_stale_threshold(fr, n, nlive, lv, lane);
                    if (lane == lv && ks != kInfF) fl = min(fl, (uint32_t)ks << lv);
                }
            }
            flim[l] = fl;
        }
        __syncwarp();

        // ---- phase 3: IP and IF of every delay (lane = delay); ps is free afterwards
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int ti = t0 + lane;
            if (ti < T) {
                int l, tp;
                if (ti < cnt0) {
                    l = 0;
                    tp = lo0 + ti;
                } else {
                    const int q = ti - cnt0;
                    l = 1 + q / DPL;
                    tp = LO + q % DPL;
                }
                const int Ll = F >> l;
                const int neff = Ll - tp;
                const double s1 = mf_pow2_neg_d(l);
                const double ipn = ps[mf_lower_bound(fr, n, (uint32_t)max(neff, 0) << l)];
                const double ifn = totD[l] - ps[mf_lower_bound(fr, n, (uint32_t)tp << l)];
                oIP[ti] = __float_as_uint(mf_scaled_div((float)(ipn * s1), neff));
                oIF[ti] = __float_as_uint(mf_scaled_div((float)(ifn * s1), neff));
            }
        }
        __syncwarp();

        // ---- phase 4: sparse levels, one pass over event pairs
        {
            const uint32_t dmax = (uint32_t)(2 * DPL + 1) << (ld - 1);
            const uint32_t top0 = (uint32_t)(lo0 + cnt0 - 1);
            const uint32_t flim0 = flim[0];
            // flattened as in multitau_warp.cu: events with partners compacted (Ic = event, Qc = its first
            // pair), owners that start inside a 32-pair chunk mark their first lane in a bit mask
            uint32_t *Qc = U;
            uint32_t *Ic = U + m.len_cap + 34;
            uint32_t npairs = 0;
            int nact = 0;
            const int topstep = n > 0 ? (1 << (31 - __clz(n))) : 1;
            for (int c0 = 0; c0 < n; c0 += 32) {
                const int i = c0 + lane;
                const uint32_t key = i < n ? min(fr[i] + dmax, (uint32_t)F) : 0u;  // partners: f_j < key
                int pos = i;
                if (__any_sync(kFullF, i + 8 < n && fr[min(i + 8, n - 1)] < key)) {
                    for (int step = topstep; step >= 8; step >>= 1) {
                        const int k2 = pos + step;
                        if (k2 < n && fr[k2] < key) pos = k2;
                    }
                }
#pragma unroll
                for (int step = 4; step >= 1; step >>= 1) {
                    const int k2 = pos + step;
                    if (k2 < n && fr[k2] < key) pos = k2;
                }
                const uint32_t mi = i < n ? (uint32_t)(pos - i) : 0u;
                uint32_t x = mi;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(kFullF, x, o);
                    if (lane >= o) x += y;
                }
                const unsigned act = __ballot_sync(kFullF, mi > 0u);
                if (mi > 0u) {
                    const int k = nact + __popc(act & ((1u << lane) - 1u));
                    Qc[k] = npairs + x - mi;
                    Ic[k] = (uint32_t)i;
                }
                nact += __popc(act);
                npairs += __shfl_sync(kFullF, x, 31);
            }
            Qc[nact + lane] = 0xffffffffu;
            if (lane < 2) Qc[nact + 32 + lane] = 0xffffffffu;
            __syncwarp();
            int kbase = 0;
            for (uint32_t p0 = 0; p0 < npairs; p0 += 32) {
                const uint32_t sl = Qc[kbase + 1 + lane] - p0;
                const unsigned bits = __reduce_or_sync(kFullF, sl < 32u ? (1u << sl) : 0u);
                const int k = kbase + __popc(bits & (0xffffffffu >> (31 - lane)));
                kbase += __popc(bits);
                const uint32_t p = p0 + lane;
                if (p < npairs) {
                    const int i = (int)Ic[k];
                    const int j = i + 1 + (int)(p - Qc[k]);
                    const uint32_t fi = fr[i], fj = fr[j];
                    const uint32_t d = fj - fi;
                    const double cc = (double)vl[i] * (double)vl[j];
                    if (d <= top0 && d >= (uint32_t)lo0 && fj < flim0) atomicAdd(&Hacc[d - lo0], cc);
                    if (d >= 2u * DPL) {
                        const int l0 = (32 - __clz((int)d)) - (LG + 1);  // d >> l0 in [dpl, 2 dpl)
                        const uint32_t b0 = (fj >> l0) - (fi >> l0) - LO;
                        if (b0 < cnts[l0] && fj < flim[l0]) atomicAdd(&Hacc[cnt0 + (l0 - 1) * DPL + b0], cc);
                        const int l1 = l0 - 1;
                        if (l1 >= 1) {
                            const uint32_t b1 = (fj >> l1) - (fi >> l1);
                            if (b1 == 2u * DPL && (uint32_t)(DPL - 1) < cnts[l1] && fj < flim[l1])
                                atomicAdd(&Hacc[cnt0 + (l1 - 1) * DPL + (DPL - 1)], cc);
                        }
                    }
                }
            }
        }
        __syncwarp();

        // ---- phase 5: dense levels, fp32 bin arrays (alias the prefix sums)
        if (ld < nl) {
            int smin_run = COMPAT ? (int)sminS[ld] : kInfF;
            for (int l = ld; l < nl; l++) {
                const int cnt_l = a.sched.count[l];
                if (cnt_l == 0) break;
                const int Ll = F >> l;
                if (l == ld) {
                    for (int t = lane; t < Ll + 64; t += 32) bh[t] = 0.0f;
                    __syncwarp();
                    // bin sums in event order: a segmented inclusive scan over the chunk, the last
                    // event of every bin adds its segment to the bin (bins straddling a chunk
                    // boundary get their parts chunk by chunk, in order)
                    const uint32_t liml = (uint32_t)Ll << l;
                    for (int c0 = 0; c0 < n; c0 += 32) {
                        const int i = c0 + lane;
                        const bool live = i < n && fr[i] < liml;
                        const uint32_t key = live ? (fr[i] >> l) : (0xfffffff0u + (uint32_t)lane);
                        float x = live ? vl[i] : 0.0f;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const float y = __shfl_up_sync(kFullF, x, o);
                            const uint32_t ko = __shfl_up_sync(kFullF, key, o);
                            if (lane >= o && ko == key) x = __fadd_rn(x, y);
                        }
                        const uint32_t kn = __shfl_down_sync(kFullF, key, 1);
                        if (live && (lane == 31 || kn != key)) bh[key] = __fadd_rn(bh[key], x);
                        __syncwarp();
                    }
                } else {
                    // B_l[t] = B_{l-1}[2t] + B_{l-1}[2t+1], in place front to back
                    for (int t0 = 0; t0 < Ll + 48; t0 += 32) {
                        const int t = t0 + lane;
                        float v = 0.0f;
                        if (t < Ll) v = __fadd_rn(bh[2 * t], bh[2 * t + 1]);
                        __syncwarp();
                        bh[t] = v;
                        __syncwarp();
                    }
                    if (COMPAT && nlive[l] < nlive[l - 1] && (int)nlive[l] < Ll) {  // all bins occupied: key >= n_l = L_l
                        // first stale slot this level leaves behind: the level-(l-1) bin of rank n_l
                        const int i = mf_select_head(fr, n, l - 1, (int)nlive[l], lane);
                        smin_run = min(smin_run, (int)(fr[i] >> (l - 1)));
                    }
                }
                int klim = Ll;
                if (COMPAT && (int)nlive[l] < n && smin_run < Ll) klim = min(Ll, mf_stale_threshold(fr, n, nlive, l, lane));
                float acc[DPL];
#pragma unroll
                for (int k = 0; k < DPL; k++) acc[k] = 0.0f;
                if (klim == Ll) {
                    for (int t0 = 0; t0 < Ll - LO; t0 += 32) {
                        const int t = t0 + lane;
                        const float x = bh[t];
#pragma unroll
                        for (int k = 0; k < DPL; k++) acc[k] = __fadd_rn(acc[k], __fmul_rn(x, bh[t + LO + k]));
                    }
                } else {
                    for (int t0 = 0; t0 < klim - LO; t0 += 32) {
                        const int t = t0 + lane;
                        const float x = bh[t];
#pragma unroll
                        for (int k = 0; k < DPL; k++)
                            if (t + LO + k < klim) acc[k] = __fadd_rn(acc[k], __fmul_rn(x, bh[t + LO + k]));
                    }
                }
                double mine = 0.0;
#pragma unroll
                for (int k = 0; k < DPL; k++) {
                    float v = acc[k];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(kFullF, v, o));
                    if (lane == k) mine = (double)v;
                }
                if (lane < cnt_l) Hacc[cnt0 + (l - 1) * DPL + lane] = mine;
            }
        }
        __syncwarp();

        // ---- phase 6: G2, one rounding to fp32 and one IEEE division per output (lane = delay)
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int ti = t0 + lane;
            if (ti < T) {
                int l, tp;
                if (ti < cnt0) {
                    l = 0;
                    tp = lo0 + ti;
                } else {
                    const int q = ti - cnt0;
                    l = 1 + q / DPL;
                    tp = LO + q % DPL;
                }
                const int neff = (F >> l) - tp;
                H[ti] = __float_as_uint(mf_scaled_div((float)(Hacc[ti] * mf_pow2_neg_d(2 * l)), neff));
            }
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- results: [3][32 rows][T] stage -> [T][R_pad], one 128-byte line per (array, delay)
    {
        float *dst[3] = {a.G2, a.IP, a.IF};
        const int64_t r0 = (int64_t)s * kSlice + lane;
#pragma unroll
        for (int arr = 0; arr < 3; arr++) {
            const uint32_t *src = outS + arr * 32 * m.pitch_t + lane * m.pitch_t;
            float *d = dst[arr] + r0;
            for (int t = warp; t < T; t += nwarps) d[(int64_t)t * a.R_pad] = __uint_as_float(src[t]);
        }
    }
    if (m.inplace && tid == 0) m.fallback[s] = 0;  // (every thread read the flag before the barrier above)
}

// Same schedule shape as the integer warp kernel; the rows must be the float store.
bool multitau_warpf_eligible(const xpcs_handle_s *h)
{
    if (h->kind != kFloat) return false;
    const int dpl = h->prm.delays_per_level;
    if (dpl != 4 && dpl != 8) return false;
    const Sched &sc = h->sched;
    if (sc.n_levels > 24 || sc.n_levels < 1) return false;
    if (sc.count[0] < 1 || sc.lo[0] != 1 || sc.first[0] != 0) return false;
    bool ended = false;
    for (int l = 1; l < sc.n_levels; l++) {
        if (sc.count[l] == 0) {
            ended = true;
            continue;
        }
        if (ended) return false;
        if (sc.lo[l] != dpl + 1 || sc.first[l] != sc.count[0] + (l - 1) * dpl || sc.count[l] > dpl) return false;
        if (l + 1 < sc.n_levels && sc.count[l + 1] > 0 && sc.count[l] != dpl) return false;
    }
    return true;
}

template <int DPL, bool COMPAT>
static int run_warpf(xpcs_handle_s *h, MtArgs &a, MfArgs &m, size_t bytes, int warps)
{
    int rc = check_cuda(h, cudaFuncSetAttribute(k_multitau_warpf<DPL, COMPAT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)bytes), "multitau_warpf smem attr");
    if (rc) return rc;
    LaunchScope ls(h, "k_multitau_warpf");
    k_multitau_warpf<DPL, COMPAT><<<h->n_slices, warps * 32, bytes, h->stream>>>(a, m);
    return XPCS_OK;
}

// flagged_only: h->d_mt_fallback holds the slices k_multitau_slicef left; they are taken here and their flags cleared
int launch_multitau_warpf(xpcs_handle_s *h, MtArgs &a, bool flagged_only)
{
    int rc = ensure(h, h->d_mt_fallback, (size_t)(h->n_slices > 0 ? h->n_slices : 1), "multitau fallback flags");
    if (rc) return rc;
    if (!flagged_only) cudaMemsetAsync(h->d_mt_fallback.p, 0, (size_t)(h->n_slices > 0 ? h->n_slices : 1), h->stream);
    if (h->n_slices == 0) return XPCS_OK;
    int smem_cap = 0;
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    MfArgs m{};
    m.fallback = h->d_mt_fallback.p;
    m.inplace = flagged_only ? 1 : 0;
    m.T = h->T;
    m.pitch_t = h->T | 1;
    // shared words: stage [3][32][pitch_t] (even), then per warp Hacc (2T) + tot (64) + tables (160)
    // + frames and values (2 len) + max(prefix sums 2 (len + 1), bins bin_cap + 64): every area even
    const size_t out_words = (size_t)3 * 32 * m.pitch_t;
    auto warp_words = [&](int len, int bins) {
        return (size_t)2 * m.T + 64 + 160 + (size_t)2 * len + std::max((size_t)bins + 64, (size_t)2 * len + 2);
    };
    // Rows beyond kMfExactLen events stay with the lane-per-row kernel, which keeps the reference's
    // sequential fp32 order: the reference's own rounding grows with the length of its chains
    // (1.7e-5 on a 3000-event row), and beyond ~1000 events sums that are more exact than the
    // reference's stop agreeing with it within the 1e-5 tolerance.
    int len_cap = h->max_row > 0 ? std::min(h->max_row, kMfExactLen) : 1;
    const size_t budget1 = ((size_t)smem_cap - 512) / 4;
    // A CTA works through its 32 rows in ceil(32 / warps) rounds: 16 warps (2 rounds) when the longest
    // row leaves them at least 2 len bins each (rows with more than bin_cap / 4 events then start their
    // dense levels one level later), else 11 warps (3 rounds) with the full 4 len, else what fits.
    int warps = 0, bin_cap = 4 * len_cap;
    if (out_words + 16 * warp_words(len_cap, 2 * len_cap) <= budget1) {
        warps = 16;
        const size_t area = ((budget1 - out_words) / 16) & ~(size_t)1;           // words per warp
        const size_t u = area - ((size_t)2 * m.T + 64 + 160 + (size_t)2 * len_cap);  // prefix sums / bins
        bin_cap = (int)(std::min((size_t)4 * len_cap, u - 64) & ~(size_t)1);
    } else if (out_words + 11 * warp_words(len_cap, 4 * len_cap) <= budget1) {
        warps = 11;
    } else {
        while (len_cap > 1 && out_words + 4 * warp_words(len_cap, 4 * len_cap) > budget1) len_cap = len_cap * 3 / 4;
        if (out_words + 4 * warp_words(len_cap, 4 * len_cap) > budget1) {  // T too large for the stage: everything falls back
            if (!flagged_only) cudaMemsetAsync(h->d_mt_fallback.p, 1, (size_t)h->n_slices, h->stream);
            return XPCS_OK;
        }
        bin_cap = 4 * len_cap;
        warps = (int)((budget1 - out_words) / warp_words(len_cap, bin_cap));
        if (warps > kMfMaxWarps) warps = kMfMaxWarps;
    }
    m.len_cap = len_cap;
    m.bin_cap = bin_cap;
    m.warp_words = (int)warp_words(len_cap, bin_cap);
    const size_t bytes = 4 * (out_words + (size_t)warps * m.warp_words);
    const bool compat = a.compat != 0;
    const int dpl = h->prm.delays_per_level;
    if (dpl == 8) rc = compat ? run_warpf<8, true>(h, a, m, bytes, warps) : run_warpf<8, false>(h, a, m, bytes, warps);
    else rc = compat ? run_warpf<4, true>(h, a, m, bytes, warps) : run_warpf<4, false>(h, a, m, bytes, warps);
    if (rc) return rc;
    return check_cuda(h, cudaGetLastError(), "k_multitau_warpf");
}

}  // namespace xpcs
