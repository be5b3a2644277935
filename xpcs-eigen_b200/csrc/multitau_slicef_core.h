// multitau_slicef_core.h -- the per-row routines of k_multitau_slicef (multitau_slicef.cu): float-valued rows.
//
// Same mapping as multitau_slice_core.h -- one LANE works on one pixel row, the warps of the CTA take different
// tasks on the same 32 rows -- for the float store (word = frame << 32 | float bits: flat-fielded, averaged,
// frame-sum-normalised or dense-source data).  The tile sits in shared memory as two planes, frames and values
// apart: fr[j * 32 + r] (0xffffffff from the end of the row on) and vl[j * 32 + r].  The frame plane is a packed
// word without a count field, so everything that depends on the frames only (level limits, live bins, first stale
// slots, the threshold key K* of SURVEY.md A.4) comes from multitau_slice_core.h, compiled here with XS_CB = 0.
//
// Mathematics as there (reference corr.cpp:315-431); arithmetic as in multitau_warpf.cu: the reference adds its
// sums sequentially in fp32, here every sum is formed in fp64 (running sums of IP / IF, the bins and the windowed
// products of the dense levels -- a product of two floats is exact in fp64) or in short fp32 chains that are
// combined in fp64 (the pair sums of the sparse levels: one private accumulator array per warp, no atomics --
// there is no native shared-memory float add), then one rounding to fp32 and one IEEE division per output.  For rows
// of up to ~1000 events that agrees with the reference to ~1e-6 relative (north_star tolerance 1e-5).
// Like its integer sibling this file compiles for the host (tests/host_mt/mt_slicef_host.cpp,
// tests/test_multitau_slicef_core.py: checked against the oracle on the CPU).
#pragma once
#include "multitau_slice_core.h"

#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

namespace xpcs {
namespace XS_NS {

static_assert(kCB == 0, "multitau_slicef_core.h needs the frame-only build of multitau_slice_core.h (XS_CB = 0)");

XS_HD double pow2_neg_d(int e)
{
    union { uint64_t u; double d; } c;
    c.u = (uint64_t)(1023 - e) << 52;
    return c.d;
}

XS_HD double fma_d(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// sum of the row's values
XS_HD double lanef_total(const float *vl, int n)
{
    double t = 0.0;
    for (int i = 0; i < n; i++) t += (double)vl[i * kS];
    return t;
}

// ---- IF of the delay slots [ta, tb): 2^-l (PS(lim_l) - PS(t' << l)) / (L_l - t'), PS(x) = values with f < x
template <int DPL>
XS_HD void lanef_if(const uint32_t *fr, const float *vl, int n, double total, const SlSched &s, int ta, int tb,
                    float *out, int64_t ostride)
{
    int q = n, pe = 0;
    double run = 0.0, dead = 0.0;
    uint32_t w = fr[0];
    uint32_t wq = n > 0 ? fr[(n - 1) * kS] : 0u;
    int l, k;
    slot_level<DPL>(s, ta, l, k);
    for (int ti = ta; ti < tb; l++, k = 0) {
        const int cnt = level_count<DPL>(s, l);
        const int Ll = s.F >> l;
        const uint32_t limf = (uint32_t)Ll << l;
        while (q > 0 && wq >= limf) {
            dead += (double)vl[(q - 1) * kS];
            q--;
            wq = q > 0 ? fr[(q - 1) * kS] : 0u;
        }
        const double totl = total - dead;
        const double s1 = pow2_neg_d(l);
        const int tp0 = l == 0 ? 1 : DPL + 1;
        for (; k < cnt && ti < tb; k++, ti++) {
            const int tp = tp0 + k;
            const uint32_t thr = (uint32_t)tp << l;
            while (w < thr) {
                run += (double)vl[pe];
                pe += kS;
                w = fr[pe];
            }
            *out = scaled_div((float)((totl - run) * s1), Ll - tp);
            out += ostride;
        }
        if (cnt == 0) break;
    }
}

// ---- IP of the delay slots [ta, tb): 2^-l PS((L_l - t') << l) / (L_l - t'); the thresholds descend with the slot
template <int DPL>
XS_HD void lanef_ip(const uint32_t *fr, const float *vl, int n, double total, const SlSched &s, int ta, int tb,
                    float *out, int64_t ostride)
{
    int q = n;
    double dead = 0.0;
    uint32_t wq = n > 0 ? fr[(n - 1) * kS] : 0u;
    int l, k;
    slot_level<DPL>(s, ta, l, k);
    for (int ti = ta; ti < tb; l++, k = 0) {
        const int cnt = level_count<DPL>(s, l);
        const int Ll = s.F >> l;
        const double s1 = pow2_neg_d(l);
        const int tp0 = l == 0 ? 1 : DPL + 1;
        for (; k < cnt && ti < tb; k++, ti++) {
            const int tp = tp0 + k;
            const int thr = (Ll - tp) << l;
            const uint32_t thrf = thr > 0 ? (uint32_t)thr : 0u;
            while (q > 0 && wq >= thrf) {
                dead += (double)vl[(q - 1) * kS];
                q--;
                wq = q > 0 ? fr[(q - 1) * kS] : 0u;
            }
            *out = scaled_div((float)((total - dead) * s1), Ll - tp);
            out += ostride;
        }
        if (cnt == 0) break;
    }
}

// ---- G2, sparse levels: pair_add / lane_pairs of multitau_slice_core.h with float products; H is the calling
// warp's own accumulator array (H[slot * 32], this lane's column), so a plain read-add-write will do
template <int DPL, bool FULL, bool CHECK>
XS_HD void pairf_add(uint32_t fi, uint32_t fj, float cc, int ld, const SlSched &s, const uint32_t *lim, float *H, float *Hl)
{
    constexpr int LG = DPL == 8 ? 3 : 2;
    const uint32_t d = fj - fi;
    if (d <= 2u * DPL) {  // rare
        if (d - 1u < (uint32_t)s.cnt0 && (!CHECK || fj < lim[0])) H[(d - 1u) * kS] += cc;
        if (d < 2u * DPL) return;
    }
    const int l0 = top_bit(d) - LG;  // d >> l0 in [dpl, 2 dpl); >= 1
    const uint32_t q0 = (fj >> l0) - (fi >> l0);
    int l = l0;
    uint32_t b = q0 - (DPL + 1);
    if (q0 == (uint32_t)DPL) {
        if (l0 < 2 || (((fj ^ fi) >> (l0 - 1)) & 1u)) return;
        l = l0 - 1;
        b = DPL - 1;
    }
    if (l >= ld) return;
    if (!FULL && b >= (uint32_t)level_count<DPL>(s, l)) return;
    if (CHECK && fj >= lim[l * kS]) return;
    Hl[((uint32_t)l * DPL + b) * kS] += cc;  // slot of (level l, bin distance dpl+1 + b)
}

template <int DPL, bool FULL>
XS_HD void lanef_pairs(const uint32_t *fr, const float *vl, int ia, int ib, int i0, int istep, int ld, const SlSched &s,
                       const uint32_t *lim, float *H)
{
    // the events i = ia + i0, ia + i0 + istep, ... below ib (ib <= n) against all their later partners
    i0 += ia;
    if (i0 >= ib) return;
    const uint32_t dmax = (uint32_t)(2 * DPL + 1) << (ld - 1);
    uint32_t limmin = lim[0];
    for (int l = 1; l < ld; l++) limmin = lim[l * kS] < limmin ? lim[l * kS] : limmin;
    int pi = i0 * kS;
    const int pend = ib * kS;
    uint32_t fi = fr[pi];
    float vi = vl[pi];
    uint32_t key = fi + dmax;  // (frames < 2^27, dmax < 2^29: no wrap; the sentinel 0xffffffff is never below it)
    int pj = pi + kS;
    float *Hl = H + (s.cnt0 - DPL) * kS;  // Hl[(l * dpl + b) * 32]: slot b of level l >= 1
    for (;;) {
        const uint32_t fj = fr[pj];
        if (fj < key) {
            const float cc = vi * vl[pj];
            pj += kS;
            if (fj < limmin) pairf_add<DPL, FULL, false>(fi, fj, cc, ld, s, lim, H, Hl);
            else pairf_add<DPL, FULL, true>(fi, fj, cc, ld, s, lim, H, Hl);
        } else {
            pi += istep * kS;
            if (pi >= pend) break;
            fi = fr[pi];
            vi = vl[pi];
            key = fi + dmax;
            pj = pi + kS;
        }
    }
}

// ---- G2, dense level l: sources t in [tb, te), targets t + dpl+1 .. t + 2dpl, bins from klim on read as zero
// (klim = L_l, or K* in compat mode: the lost targets are a suffix).  te - tb must be a multiple of 2dpl+1 unless
// te >= L_l.  The bins are formed on the fly (fp32 sum of the few events of a bin, as the reference's own bins are)
// while a register window of 2dpl+1 bins slides over t; the windowed products are added up in fp64 (acc[] is added
// to).  Only the last windows before klim ask whether a bin lies beyond it.
template <int DPL>
XS_HD void lanef_dense(const uint32_t *fr, const float *vl, int n, int l, int tb, int te, int klim, double (&acc)[DPL])
{
    constexpr int W = 2 * DPL + 1;
    int pe = (tb > 0 ? lane_lower_bound(fr, n, (uint32_t)tb << l) : 0) * kS;
    uint32_t wk = fr[pe] >> l;  // key of the next event (the sentinel's is beyond every bin)
    auto fetch = [&](int key) -> double {
        float v = 0.0f;
        while (wk == (uint32_t)key) {
            v += vl[pe];
            pe += kS;
            wk = fr[pe] >> l;
        }
        return (double)v;
    };
    double win[W];
#pragma unroll
    for (int k = 0; k < W; k++) win[k] = tb + k < klim ? fetch(tb + k) : 0.0;
    int t0 = tb;
    for (; t0 < te && t0 + 2 * W <= klim; t0 += W) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            const double src = win[u];
#pragma unroll
            for (int d = 0; d < DPL; d++) acc[d] = fma_d(src, win[(u + DPL + 1 + d) % W], acc[d]);
            win[u] = fetch(t0 + u + W);
        }
    }
    for (; t0 < te; t0 += W) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            const double src = win[u];
#pragma unroll
            for (int d = 0; d < DPL; d++) acc[d] = fma_d(src, win[(u + DPL + 1 + d) % W], acc[d]);
            win[u] = t0 + u + W < klim ? fetch(t0 + u + W) : 0.0;
        }
    }
}

// ---- pieces of the dense levels ld .. lastl: level l is cut into max(1, ceil(L_l / target)) pieces of `target` bins,
// target = a multiple of 2dpl+1 close to (all dense bins) / nd.  Every piece is a task of its own and leaves its sums in
// its own slot, so the order in which they are added up is fixed.
template <int DPL>
XS_HD int dense_target(const SlSched &s, int ld, int nd)
{
    constexpr int W = 2 * DPL + 1;
    int bins = 0;
    for (int l = ld; l <= s.lastl; l++) bins += s.F >> l;
    int target = ((bins + nd - 1) / nd + W - 1) / W * W;
    return target < W ? W : target;
}

XS_HD int dense_pieces(const SlSched &s, int l, int target)
{
    const int np = ((s.F >> l) + target - 1) / target;
    return np < 1 ? 1 : np;
}

// G2 of delay slot ti from its numerator
template <int DPL>
XS_HD float g2f_value(double num, int ti, const SlSched &s)
{
    int l, tp;
    if (ti < s.cnt0) {
        l = 0;
        tp = 1 + ti;
    } else {
        const int q = ti - s.cnt0;
        l = 1 + q / DPL;
        tp = DPL + 1 + q % DPL;
    }
    return scaled_div((float)(num * pow2_neg_d(2 * l)), (s.F >> l) - tp);
}

}  // namespace XS_NS
}  // namespace xpcs
