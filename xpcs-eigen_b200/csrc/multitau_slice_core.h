// multitau_slice_core.h -- the per-row routines of k_multitau_slice (multitau_slice.cu).
//
// One LANE works on one pixel row; the 32 rows of a slice sit in shared memory exactly as in the
// store (word j of row r at ev[j * 32 + r]: bank == lane, conflict free wherever the lanes are),
// and the warps of the CTA take different TASKS on the same 32 rows.  Nothing in this file talks
// to another lane, so every routine is plain sequential code over one column -- which is also why
// the very same source compiles for the host: tests/host_mt builds it with g++ and checks it bit
// for bit against the oracle on the CPU (tests/test_multitau_slice_core.py) before it ever
// reaches a GPU.
//
// Mathematics (reference corr.cpp:315-431, restated in tests/multitau_model.py): every quantity
// is a function of the row's level-0 events (f_i, c_i), word = f << 12 | c, ascending in f:
//   L_l = F >> l, lim_l = L_l << l; an event is live at level l iff f < lim_l   (corr.cpp:349-390)
//   IP(l,t') = PS((L_l - t') << l), IF(l,t') = PS(lim_l) - PS(t' << l), PS(x) = counts with f < x
//       (corr.cpp:403, 414-416): the IF thresholds ascend with the delay index, the IP thresholds
//       descend, so each is ONE walk over the row with a running sum;
//   G2, sparse levels (l < ld): one walk over the event pairs i < j with f_j - f_i < (2dpl+1) << (ld-1);
//       the distance decides the one or two (level, delay) slots of a pair   (corr.cpp:397-411)
//   G2, dense levels (l >= ld): the bins B_l[t] = sum of the counts with f >> l == t are formed on
//       the fly while a register window of 2dpl+1 bins slides over t -- no bin array anywhere;
//   one IEEE division per output.
// Sentinel: ev[j * 32] = 0xffffffff for n <= j <= len (stops the forward walks).
// XPCS_COMPAT_STALE_TAIL (corr.cpp:406): the threshold key K* of multitau.cu, from
// rank/select over the level-0 events; see lane_stale_threshold.
#pragma once
#include <stdint.h>

#if defined(XS_HD)
// (multitau_slice.cu: device only, with its own XS_WALK)
#elif defined(__CUDACC__)
#define XS_HD __host__ __device__ __forceinline__
#define XS_HD_CALL __host__ __device__ __noinline__
#else
#define XS_HD inline
#define XS_HD_CALL inline   // on the device: rare or long routines, one copy, called (the tasks share one instruction cache)
#endif

#if defined(__CUDA_ARCH__)
#define XS_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define XS_ADD(ptr, v) (*(ptr) += (v))
#endif

// The float-row kernel (multitau_slicef.cu) compiles this file a second time, in its own namespace and with
// XS_CB = 0: its frame plane is a packed word without a count field, so every routine that looks at frames only
// (bin heads, live counts, first stale slots, the threshold key K*) serves both kernels from one source.
#ifndef XS_NS
#define XS_NS sl
#endif
#ifndef XS_CB
#define XS_CB 12
#endif

namespace xpcs {
namespace XS_NS {

constexpr int kS = 32;                 // rows per slice = stride of a row's column
constexpr int kCB = XS_CB;             // count bits of the packed word (== kCountBits; 0: frames only)
constexpr uint32_t kCMask = (1u << kCB) - 1u;
constexpr uint32_t kSent = 0xffffffffu;
constexpr int kInfKey = 0x7fffffff;
constexpr int kMlRows = 33;            // merge-level histogram rows (bit lengths 0..32)

// launch-uniform schedule: level 0 has the delays 1..cnt0, level l in 1..lastl the level-local delays
// dpl+1..dpl+count (count = dpl below lastl, cnt_last at lastl)
struct SlSched {
    int F, nl, T, cnt0, lastl, cnt_last;
};

XS_HD int bitlength(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return 32 - __clz((int)x);
#else
    return x ? 32 - __builtin_clz(x) : 0;
#endif
}

// index of the highest set bit (x != 0)
XS_HD int top_bit(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
#else
    return 31 - __builtin_clz(x);
#endif
}

XS_HD float pow2_neg(int e)
{
    union { uint32_t u; float f; } c;
    c.u = (uint32_t)(127 - e) << 23;
    return c.f;
}

XS_HD float scaled_div(float num, int neff)
{
#if defined(__CUDA_ARCH__)
    return neff > 0 ? __fdiv_rn(num, (float)neff) : num;
#else
    return neff > 0 ? num / (float)neff : num;
#endif
}

template <int DPL>
XS_HD int level_count(const SlSched &s, int l)
{
    return l == 0 ? s.cnt0 : (l < s.lastl ? DPL : (l == s.lastl ? s.cnt_last : 0));
}

// first index in [0, n) whose word is >= keyw (n if none)
XS_HD int lane_lower_bound(const uint32_t *ev, int n, uint32_t keyw)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ev[mid * kS] < keyw) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// level and position inside the level of delay slot ti
template <int DPL>
XS_HD void slot_level(const SlSched &s, int ti, int &l, int &k)
{
    if (ti < s.cnt0) {
        l = 0;
        k = ti;
    } else {
        const int q = ti - s.cnt0;
        l = 1 + q / DPL;
        k = q % DPL;
    }
}

// ---- IF of the delay slots [ta, tb): PS(lim_l) - PS(t' << l).  Forward walk for PS(t' << l) (the thresholds
// ascend with the slot, so a range simply starts with a longer first step); the events beyond lim_l are the last
// few of the row and leave through a second, backward pointer as the levels go up.  out points at slot ta.
template <int DPL>
XS_HD void lane_if(const uint32_t *ev, int n, uint32_t total, const SlSched &s, int ta, int tb, float *out, int64_t ostride)
{
    int q = n;
    const uint32_t *pe = ev;
    uint32_t run = 0, dead = 0;
    uint32_t w = ev[0];
    uint32_t wq = n > 0 ? ev[(n - 1) * kS] : 0u;
    int l, k;
    slot_level<DPL>(s, ta, l, k);
    for (int ti = ta; ti < tb; l++, k = 0) {
        const int cnt = level_count<DPL>(s, l);
        const int Ll = s.F >> l;
        const uint32_t limw = ((uint32_t)Ll << l) << kCB;
        while (q > 0 && wq >= limw) {
            dead += wq & kCMask;
            q--;
            wq = q > 0 ? ev[(q - 1) * kS] : 0u;
        }
        const uint32_t totl = total - dead;
        const float s1 = pow2_neg(l);
        const int tp0 = l == 0 ? 1 : DPL + 1;
        for (; k < cnt && ti < tb; k++, ti++) {
            const int tp = tp0 + k;
            const uint32_t thrw = ((uint32_t)tp << l) << kCB;
            while (w < thrw) {
                run += w & kCMask;
                pe += kS;
                w = *pe;
            }
            *out = scaled_div((float)(totl - run) * s1, Ll - tp);
            out += ostride;
        }
        if (cnt == 0) break;  // (not reached with a regular schedule)
    }
}

// ---- IP of the delay slots [ta, tb): PS((L_l - t') << l); the thresholds descend with the slot: one backward walk
template <int DPL>
XS_HD void lane_ip(const uint32_t *ev, int n, uint32_t total, const SlSched &s, int ta, int tb, float *out, int64_t ostride)
{
    int q = n;
    uint32_t dead = 0;
    uint32_t wq = n > 0 ? ev[(n - 1) * kS] : 0u;
    int l, k;
    slot_level<DPL>(s, ta, l, k);
    for (int ti = ta; ti < tb; l++, k = 0) {
        const int cnt = level_count<DPL>(s, l);
        const int Ll = s.F >> l;
        const float s1 = pow2_neg(l);
        const int tp0 = l == 0 ? 1 : DPL + 1;
        for (; k < cnt && ti < tb; k++, ti++) {
            const int tp = tp0 + k;
            const int thr = (Ll - tp) << l;
            const uint32_t thrw = thr > 0 ? (uint32_t)thr << kCB : 0u;
            while (q > 0 && wq >= thrw) {
                dead += wq & kCMask;
                q--;
                wq = q > 0 ? ev[(q - 1) * kS] : 0u;
            }
            *out = scaled_div((float)(total - dead) * s1, Ll - tp);
            out += ostride;
        }
        if (cnt == 0) break;
    }
}

// ---- G2, sparse levels: the pairs (i, j), i = ia + i0, ia + i0 + istep, ... < ib, j > i, f_j - f_i < dmax.  ONE loop whose
// body either takes the next partner or moves on to the next i, so that a lane's trip count is its own
// (events + pairs) and the warp's the maximum of those -- not the sum of per-event maxima.
// lim[l * 32]: frame limit of level l (lim_l, or K* << l in compat mode); H[slot * 32]: numerators.
// A pair whose later event lies below every level's limit (all but the last few events of a row) skips the
// per-level limit look-ups.
// A pair at frame distance d >= 2 dpl sits in exactly one (level, delay) slot, or in none: with l0 such that
// d >> l0 is in [dpl, 2 dpl), the bin distance at level l0 is q0 = (f_j >> l0) - (f_i >> l0), dpl <= q0 <= 2 dpl.
// q0 > dpl: slot (l0, q0).  q0 == dpl: nothing at level l0, and at level l0 - 1 the bin distance 2 q0 + (bit_j - bit_i)
// is the last delay 2 dpl iff the two frames agree in bit l0 - 1.  Every other level sees a distance outside
// dpl+1 .. 2 dpl.  (Distances up to 2 dpl are the level-0 delays; d == 2 dpl takes part in both.)
template <int DPL, bool FULL, bool CHECK>
XS_HD void pair_add(uint32_t fi, uint32_t fj, uint32_t cc, int ld, const SlSched &s, const uint32_t *lim, uint32_t *H,
                    uint32_t *Hl)
{
    // FULL: every sparse level l in 1..ld-1 has all its dpl delays (ld - 1 < lastl), no need to ask the schedule
    constexpr int LG = DPL == 8 ? 3 : 2;
    const uint32_t d = fj - fi;
    if (d <= 2u * DPL) {  // rare
        if (d - 1u < (uint32_t)s.cnt0 && (!CHECK || fj < lim[0])) XS_ADD(&H[(d - 1u) * kS], cc);
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
    // slot of (level l, bin distance dpl+1 + b): Hl[(l * dpl + b) * 32]
    XS_ADD(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(Hl) + ((uint32_t)l << (LG + 7)) + (b << 7)), cc);
}

template <int DPL, bool FULL>
XS_HD void lane_pairs(const uint32_t *ev, int ia, int ib, int i0, int istep, int ld, const SlSched &s,
                      const uint32_t *lim, uint32_t *H)
{
    // the events i = ia + i0, ia + i0 + istep, ... below ib (ib <= n) against all their later partners
    i0 += ia;
    if (i0 >= ib) return;
    const uint32_t dmax = (uint32_t)(2 * DPL + 1) << (ld - 1);
    uint32_t limmin = lim[0];
    for (int l = 1; l < ld; l++) limmin = lim[l * kS] < limmin ? lim[l * kS] : limmin;
    const uint32_t limminw = limmin << kCB;
    const uint32_t *pi = ev + i0 * kS, *pend = ev + ib * kS;
    uint32_t wi = *pi;
    uint32_t fi = wi >> kCB, ci = wi & kCMask;
    uint32_t fend = fi + dmax;
    uint32_t keyw = (fend < 0xfffffu ? fend : 0xfffffu) << kCB;
    const uint32_t *pj = pi + kS;
    uint32_t *Hl = H + (s.cnt0 - DPL) * kS;  // Hl[(l * dpl + b) * 32]: slot b of level l >= 1
    for (;;) {
        const uint32_t wj = *pj;
        if (wj < keyw) {
            pj += kS;
            const uint32_t fj = wj >> kCB;
            const uint32_t cc = ci * (wj & kCMask);
            if (wj < limminw) pair_add<DPL, FULL, false>(fi, fj, cc, ld, s, lim, H, Hl);
            else pair_add<DPL, FULL, true>(fi, fj, cc, ld, s, lim, H, Hl);
        } else {
            pi += istep * kS;
            if (pi >= pend) break;
            wi = *pi;
            fi = wi >> kCB;
            ci = wi & kCMask;
            fend = fi + dmax;
            keyw = (fend < 0xfffffu ? fend : 0xfffffu) << kCB;
            pj = pi + kS;
        }
    }
}

// ---- G2, dense level l: sources t in [tb, te), targets t + dpl+1 .. t + 2dpl, bins from klim on read as
// zero (klim = L_l, or K* in compat mode: the lost targets are a suffix).  te - tb must be a multiple of
// 2dpl+1 unless te >= L_l (sources past the end have no live target).  acc[] is added to.
template <int DPL>
XS_HD void lane_dense(const uint32_t *ev, int n, int l, int tb, int te, int klim, uint32_t (&acc)[DPL])
{
    constexpr int W = 2 * DPL + 1;
    const int sh = kCB + l;
    const uint32_t *pe = ev + (tb > 0 ? lane_lower_bound(ev, n, (uint32_t)tb << sh) : 0) * kS;
    uint32_t w = *pe;
    auto fetch = [&](int key) -> uint32_t {
        uint32_t v = 0;
        if (key < klim) {
            while ((int)(w >> sh) == key) {
                v += w & kCMask;
                pe += kS;
                w = *pe;
            }
        }
        return v;
    };
    uint32_t win[W];
#pragma unroll
    for (int k = 0; k < W; k++) win[k] = fetch(tb + k);
    for (int t0 = tb; t0 < te; t0 += W) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            const uint32_t src = win[u];
#pragma unroll
            for (int d = 0; d < DPL; d++) acc[d] += src * win[(u + DPL + 1 + d) % W];
            win[u] = fetch(t0 + u + W);
        }
    }
}

XS_HD_CALL int lane_stale_threshold(const uint32_t *ev, int n, const uint32_t *nlive, int level);
// K* of this lane's row if `fire`, else kInfKey.  `level` is the same for all lanes of a warp and the call sites are
// reached by all of them together, so the device build can give the rare walk to the whole warp (multitau_slice.cu).
#ifndef XS_WALK
#define XS_WALK(fire, ev, n, nlive, level) ((fire) ? lane_stale_threshold((ev), (n), (nlive), (level)) : kInfKey)
#endif

// ---- G2, dense levels, rows whose counts sum to <= 255: 8-bit bin arrays B[t * 32] (one byte per row and bin).
// Unlike the on-the-fly walk above, nothing here depends on where a lane's events lie: all lanes run the
// same trip counts.

// bins of level l from the events.  COMPAT: cand = key of the occupied level-(l-1) bin of rank `want` (0-based;
// the first stale slot level l leaves behind), kInfKey if there is none; TWO: the same one level up
// (cand2: occupied level-(l-2) bin of rank want2), for the task that starts two levels beyond the first dense one
template <bool COMPAT, bool TWO>
XS_HD void lane_bins8_build(const uint32_t *ev, int l, int F, uint8_t *B, int want, int &cand, int want2, int &cand2)
{
    const uint32_t limw = ((uint32_t)(F >> l) << l) << kCB;
    const int sh = kCB + l - 1;
    const uint32_t *pe = ev;
    uint32_t w = *pe;
    int occ = -1, occ2 = -1;
    uint32_t prev = 0xffffffffu, prev2 = 0xffffffffu;
    cand = kInfKey;
    cand2 = kInfKey;
    while (w < limw) {
        const uint32_t key1 = w >> sh;  // level l - 1
        B[(key1 >> 1) * kS] += (uint8_t)(w & kCMask);
        if (COMPAT && key1 != prev) {
            prev = key1;
            if (++occ == want) cand = (int)key1;
        }
        if (COMPAT && TWO) {
            const uint32_t key2 = w >> (sh - 1);  // level l - 2
            if (key2 != prev2) {
                prev2 = key2;
                if (++occ2 == want2) cand2 = (int)key2;
            }
        }
        pe += kS;
        w = *pe;
    }
}

// B_{l+1}[t] = B_l[2t] + B_l[2t+1] in place, t < Lnext; COMPAT: cand as above from the level-l bins
template <bool COMPAT>
XS_HD void lane_bins8_halve(uint8_t *B, int Lnext, int want, int &cand)
{
    int occ = -1;
    cand = kInfKey;
    for (int t = 0; t < Lnext; t++) {
        const uint32_t a = B[(2 * t) * kS], b = B[(2 * t + 1) * kS];
        if (COMPAT) {
            if (a != 0u && ++occ == want) cand = 2 * t;
            if (b != 0u && ++occ == want) cand = 2 * t + 1;
        }
        B[t * kS] = (uint8_t)(a + b);
    }
}

// acc[d] = sum over t of B[t] * B[t + dpl+1 + d], t + dpl+1 + d < L: a register window of 2dpl+1 bins slides over
// t; only the last two windows ask whether they have run past the end of the level
template <int DPL>
XS_HD void lane_bins8_mac(const uint8_t *B, int L, uint32_t (&acc)[DPL])
{
    constexpr int W = 2 * DPL + 1;
    uint32_t win[W];
#pragma unroll
    for (int k = 0; k < W; k++) win[k] = k < L ? B[k * kS] : 0u;
    const uint8_t *pb = B + W * kS;
    int t0 = 0;
    for (; t0 + 2 * W <= L; t0 += W, pb += W * kS) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            const uint32_t src = win[u];
#pragma unroll
            for (int d = 0; d < DPL; d++) acc[d] += src * win[(u + DPL + 1 + d) % W];
            win[u] = pb[u * kS];
        }
    }
    for (; t0 < L; t0 += W, pb += W * kS) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            const uint32_t src = win[u];
#pragma unroll
            for (int d = 0; d < DPL; d++) acc[d] += src * win[(u + DPL + 1 + d) % W];
            win[u] = t0 + u + W < L ? pb[u * kS] : 0u;
        }
    }
}

// compat: take the pairs whose target bin is >= klim out again (the lost targets are a suffix; rare rows only)
template <int DPL>
XS_HD void lane_bins8_fix(const uint8_t *B, int L, int klim, uint32_t (&acc)[DPL])
{
    for (int tt = klim < 0 ? 0 : klim; tt < L; tt++) {
        const uint32_t x = B[tt * kS];
        if (x == 0u) continue;
#pragma unroll
        for (int d = 0; d < DPL; d++) {
            const int src = tt - (DPL + 1) - d;
            if (src >= 0) acc[d] -= (uint32_t)B[src * kS] * x;
        }
    }
}

// one dense level, bins in B: numerators, the compat correction, one division per delay
template <int DPL>
XS_HD_CALL void lane_dense8_level(const uint8_t *B, int l, int klim, const SlSched &s, float *g2, int64_t ostride)
{
    const int L = s.F >> l;
    uint32_t acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; d++) acc[d] = 0u;
    lane_bins8_mac<DPL>(B, L, acc);
    if (klim < L) lane_bins8_fix<DPL>(B, L, klim, acc);
    const int cnt = level_count<DPL>(s, l);
    const int slot0 = s.cnt0 + (l - 1) * DPL;
    const float s2 = pow2_neg(2 * l);
#pragma unroll
    for (int d = 0; d < DPL; d++)
        if (d < cnt) g2[(slot0 + d) * ostride] = scaled_div((float)acc[d] * s2, L - (DPL + 1 + d));
}

// Three tasks share the dense levels, each with its own bin array (zeroed by the caller) filled from the events:
//   which = 0   level ld                 (key limit from the tables)
//   which = 1   level ld + 1
//   which = 2   levels ld + 2 .. lastl   (bins halved in place from level to level)
// In compat mode the exact first stale slot of every level beyond ld comes out of the same passes (smin0: the
// smallest one up to level ld, from the tables), and the threshold key K* is looked for only where it can matter.
template <int DPL, bool COMPAT>
XS_HD void lane_dense8(int which, const uint32_t *ev, int n, int ld, const SlSched &s, uint8_t *B, const uint32_t *lim,
                       const uint32_t *nlive, int smin0, float *g2, int64_t ostride)
{
    int cand, cand2;
    if (which == 0) {
        lane_bins8_build<false, false>(ev, ld, s.F, B, -2, cand, -2, cand2);
        lane_dense8_level<DPL>(B, ld, (int)lim[ld * kS], s, g2, ostride);
        return;
    }
    int smin = smin0;
    // compat: first stale slot left by level l (from the level-(l-1) bins) joins the running minimum; K* if it can matter
    auto key_limit = [&](int l, int cnd) -> int {
        const int L = s.F >> l;
        if (!COMPAT) return L;
        const int nv = (int)nlive[l * kS];
        if (nv < (int)nlive[(l - 1) * kS] && nv < L) smin = cnd < smin ? cnd : smin;
        const int ks = XS_WALK(nv < n && smin < L, ev, n, nlive, l);
        return ks < L ? ks : L;
    };
    if (which == 1) {
        const int l = ld + 1;
        lane_bins8_build<COMPAT, false>(ev, l, s.F, B, COMPAT ? (int)nlive[l * kS] : -2, cand, -2, cand2);
        lane_dense8_level<DPL>(B, l, key_limit(l, cand), s, g2, ostride);
        return;
    }
    for (int l = ld + 2; l <= s.lastl; l++) {
        if (l == ld + 2) {
            lane_bins8_build<COMPAT, true>(ev, l, s.F, B, COMPAT ? (int)nlive[l * kS] : -2, cand,
                                           COMPAT ? (int)nlive[(l - 1) * kS] : -2, cand2);
            if (COMPAT) {  // level ld + 1 belongs to another task, but its first stale slot counts here too
                const int nv = (int)nlive[(l - 1) * kS];
                if (nv < (int)nlive[(l - 2) * kS] && nv < (s.F >> (l - 1))) smin = cand2 < smin ? cand2 : smin;
            }
        } else lane_bins8_halve<COMPAT>(B, s.F >> l, COMPAT ? (int)nlive[l * kS] : -2, cand);
        lane_dense8_level<DPL>(B, l, key_limit(l, cand), s, g2, ostride);
    }
}

// ---- G2 output of one delay slot
template <int DPL>
XS_HD float g2_value(uint32_t num, int ti, const SlSched &s)
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
    return scaled_div((float)num * pow2_neg(2 * l), (s.F >> l) - tp);
}

// =====================================================================================================
// compat tables.  The reference searches its un-shrunk vector: slots [0, n_l) hold the live
// level-l keys, slots [n_l, n_0) what earlier levels left there: V_l[p] = A_m[p], m = max{ j <= l : n_j > p },
// A_m[p] = key of the p-th bin head of level m.  All of it follows from the merge levels
// ml_i = bitlength(f_i ^ f_{i-1}) (event i starts a bin at level l iff ml_i > l).

// histogram of the merge levels of the events [max(i0, 1), i1)
XS_HD void lane_mlhist(const uint32_t *ev, int i0, int i1, uint32_t *cntml)
{
    for (int i = i0 < 1 ? 1 : i0; i < i1; i++) {
        const int m = bitlength((ev[i * kS] ^ ev[(i - 1) * kS]) >> kCB);
        XS_ADD(&cntml[m * kS], 1u);
    }
}

// index of the p-th (0-based) event that starts a bin at `level`
XS_HD int lane_select_head(const uint32_t *ev, int n, int level, int p)
{
    int k = -1;
    for (int i = 0; i < n; i++) {
        const bool head = i == 0 || ((((ev[i * kS] ^ ev[(i - 1) * kS]) >> kCB) >> level) != 0u);
        if (head && ++k == p) return i;
    }
    return n - 1;
}

// live bins of level l >= 1 and the smallest key its compaction leaves in the first stale slot: a lower
// bound f[n_l] >> (l-1) up to the first dense level (tight there: few events have merged), the exact
// A_{l-1}[n_l] beyond it (where the bound would fire on every row)
XS_HD void lane_level_base(const uint32_t *ev, int n, int l, int ld, int F, const uint32_t *cntml,
                           uint32_t *nlive, uint32_t *sbx, bool exact_beyond_ld)
{
    int ab = 0;  // events that do not start a bin at level l
    for (int m = 1; m <= l; m++) ab += (int)cntml[m * kS];
    const int abp = ab - (int)cntml[l * kS];
    const uint32_t fl = n > 0 ? ev[(n - 1) * kS] >> kCB : 0u;
    const int dropped = (n > 0 && (int)(fl >> l) >= (F >> l)) ? 1 : 0;
    const int droppedp = (n > 0 && (int)(fl >> (l - 1)) >= (F >> (l - 1))) ? 1 : 0;
    const int nv = n - ab - dropped, nvp = n - abp - droppedp;
    nlive[l * kS] = (uint32_t)nv;
    int sb = kInfKey;
    if (nv < nvp) {
        if (l <= ld) sb = (int)((ev[nv * kS] >> kCB) >> (l - 1));
        else if (exact_beyond_ld && nv < (F >> l)) {  // (the 8-bit bin path finds these while it halves its bins)
            const int i = lane_select_head(ev, n, l - 1, nv);
            sb = (int)((ev[i * kS] >> kCB) >> (l - 1));
        }
    }
    sbx[l * kS] = (uint32_t)sb;
}

// value the reference sees in slot p of its vector at `level`
XS_HD int lane_key_at(const uint32_t *ev, int n, const uint32_t *nlive, int level, int p)
{
    int lv = level;
    if (p >= (int)nlive[level * kS]) {
        lv = level - 1;
        while (lv > 0 && (int)nlive[lv * kS] <= p) lv--;
    }
    const int i = lane_select_head(ev, n, lv, p);
    return (int)((ev[i * kS] >> kCB) >> lv);
}

// threshold key K*: targets with key >= K* are never found by the reference's search (the walk of
// multitau.cu: stale_tail_threshold, along the live/stale boundary of the implicit search tree)
XS_HD_CALL int lane_stale_threshold(const uint32_t *ev, int n, const uint32_t *nlive, int level)
{
    const int nl = (int)nlive[level * kS];
    int first = 0, len = n;
    int curmin = kInfKey;
    while (len > 0) {
        const int half = len >> 1;
        const int mid = first + half;
        if (mid >= nl) {
            const int k = lane_key_at(ev, n, nlive, level, mid);
            curmin = k < curmin ? k : curmin;
            len = half;
        } else {
            if (curmin != kInfKey && lane_key_at(ev, n, nlive, level, mid) > curmin) {
                const int k1 = lane_key_at(ev, n, nlive, level, first);
                const int j = lane_lower_bound(ev, n, ((uint32_t)(curmin + 1) << level) << kCB);
                const int k2 = j < n ? (int)((ev[j * kS] >> kCB) >> level) : kInfKey;
                return k1 > k2 ? k1 : k2;
            }
            first = mid + 1;
            len = len - half - 1;
        }
    }
    return kInfKey;
}

// frame limit (sparse levels, l < ld) or key limit (dense levels) of level l >= 1 after the stale-tail rule
XS_HD uint32_t lane_level_limit(const uint32_t *ev, int n, int l, int ld, int F, const uint32_t *nlive,
                                const uint32_t *sbx)
{
    const int Ll = F >> l;
    uint32_t out = l < ld ? (uint32_t)Ll << l : (uint32_t)Ll;
    int smin = kInfKey;
    for (int j = 1; j <= l; j++) {
        const int v = (int)sbx[j * kS];
        smin = v < smin ? v : smin;
    }
    const int ks = XS_WALK((int)nlive[l * kS] < n && smin < Ll, ev, n, nlive, l);
    if (ks != kInfKey) {
        const uint32_t lim = l < ld ? (uint32_t)ks << l : (uint32_t)ks;
        out = lim < out ? lim : out;
    }
    return out;
}

}  // namespace XS_NS
}  // namespace xpcs
