// multitau_warp.cu -- multi-tau correlator, one WARP per pixel row (integer photon counts).
//
// Replaces Corr::multiTau2 (reference corr.cpp:315-431) for the packed store
// (word = frame << 12 | count).  One CTA owns one slice of 32 rows -- or half of one, 16 rows, when
// that makes three CTAs fit an SM: the tile is read once with coalesced 128-byte (64-byte) lines and
// transposed into shared memory (row-major, odd pitch), then each warp works through its rows with
// lanes over events / bins / delays, and the rows x T x 3 results leave through a shared-memory stage
// as full 128-byte (64-byte) lines.
//
// Nothing is compacted level by level.  Every quantity is a function of the level-0 events
// (f_i, c_i) of the row (tests/multitau_model.py restates this on the CPU and is checked bit
// for bit against the oracle):
//   L_l = F >> l, lim_l = L_l << l;  an event is live at level l iff f < lim_l  (corr.cpp:349-390)
//   IP(l,t') = PS((L_l - t') << l),  IF(l,t') = PS(lim_l) - PS(t' << l)           (corr.cpp:403,414-416)
//       with PS(x) = sum of the counts of the events with f < x: the thresholds are monotone in the delay
//       index, so every event adds its count to the first delay it matters for (closed form from bit
//       lengths) and running sums over the delays give all PS at once; PS(lim_l) likewise from the level
//       bitlength(f ^ F) at which an event dies
//   G2, sparse levels (l < ld): ONE pass over event pairs i < j with d = f_j - f_i < (2dpl+1) << (ld-1);
//       a pair lands at level 0 when d <= 2dpl, and for d >= 2dpl only at the levels
//       l0 = bitlength(d) - log2(dpl) - 1 (bin distance b = (f_j >> l0) - (f_i >> l0) in dpl..2dpl)
//       and l0 - 1 (only b == 2dpl);                                               (corr.cpp:397-411)
//   G2, dense levels (l >= ld, the first level with L_l <= 4 n): 16-bit bin arrays B_l[t]
//       (B_{l+1}[t] = B_l[2t] + B_l[2t+1]) and windowed products, lanes over t.
//   All numerators are exact integers (< 2^32 because the row's counts sum to < 2^16; rows
//   beyond that go to the lane-per-row kernel), one IEEE division per output.
// XPCS_COMPAT_STALE_TAIL: the reference's binary search runs over the stale
// tail of its un-shrunk vector.  V_l[p] = A_m[p], m = max{ j <= l : n_j > p }, is a pure
// function of the level-0 events: the live counts n_l come from the merge levels
// bitlength(f_i ^ f_{i-1}); the first stale slot of each level (its key is what can steer a
// search wrong) is bounded from below on the sparse levels and tracked exactly on the dense
// ones; only when it drops below L_l is the boundary walk of multitau.cu replayed, with
// rank/select over the events instead of a materialised array.
#include <algorithm>
#include <cstdlib>

#include "internal.h"

namespace xpcs {

constexpr int kMwWarps = 16;
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kInf = 0x7fffffff;
constexpr int kMwTables = 6 * 32;  // per-warp tables: tot, flim, cnts, cntml, nlive, smin

struct MwArgs {
    unsigned char *fallback;   // [n_slices]
    int len_cap;               // longest slice handled here
    int pitch_e;               // odd
    int pitch_t;               // odd, >= T
    int warp_words;            // per-warp shared words
    int T;
    int lastl, cnt_last;       // last level >= 1 with delays (0 = none) and its delay count
    int parts;                 // CTAs per slice (1, 2 or 4): a CTA stages 32 / parts rows
    int ld_factor;             // dense levels start where L_l <= ld_factor * n (4: measured best of 1..4 on C3 and C1 -- 5.64 / 4.37 / 4.18 / 4.08 ms; at most 4, the bin buffer)
};

__device__ __forceinline__ float mw_pow2_neg(int e) { return __int_as_float((127 - e) << 23); }
__device__ __forceinline__ float mw_scaled_div(float num, int neff)
{
    return neff > 0 ? __fdiv_rn(num, (float)neff) : num;
}

// first index in [0, n) whose word is >= key (words ascend with the frame)
__device__ __forceinline__ int mw_lower_bound(const uint32_t *ev, int n, uint32_t key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ev[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// event index of the p-th (0-based) event that starts a bin at `level`; warp-uniform
__device__ __forceinline__ int mw_select_head(const uint32_t *ev, int n, int level, int p, int lane)
{
    int base = 0;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int i = c0 + lane;
        bool head = false;
        if (i < n) head = (i == 0) || ((((ev[i] ^ ev[i - 1]) >> kCountBits) >> level) != 0u);
        const unsigned mk = __ballot_sync(kFull, head);
        const int c = __popc(mk);
        if (p < base + c) return c0 + (int)__fns(mk, 0, p - base + 1);
        base += c;
    }
    return n - 1;
}

// value the reference sees in slot p of its vector at `level`; warp-uniform
__device__ __forceinline__ int mw_key_at(const uint32_t *ev, int n, const uint32_t *nlive, int level, int p, int lane)
{
    int lv = level;
    if (p >= (int)nlive[level]) {
        lv = level - 1;
        while (lv > 0 && (int)nlive[lv] <= p) lv--;
    }
    const int i = mw_select_head(ev, n, lv, p, lane);
    return (int)((ev[i] >> kCountBits) >> lv);
}

// threshold key K*: targets with key >= K* are never found by the reference's search
__device__ __forceinline__ int mw_stale_threshold(const uint32_t *ev, int n, const uint32_t *nlive, int level, int lane)
{
    const int nl = (int)nlive[level];
    int first = 0, len = n;
    int curmin = kInf;
    while (len > 0) {
        const int half = len >> 1;
        const int mid = first + half;
        if (mid >= nl) {
            curmin = min(curmin, mw_key_at(ev, n, nlive, level, mid, lane));
            len = half;
        } else {
            if (curmin != kInf && mw_key_at(ev, n, nlive, level, mid, lane) > curmin) {
                const int k1 = mw_key_at(ev, n, nlive, level, first, lane);
                const int j = mw_lower_bound(ev, n, ((uint32_t)(curmin + 1) << level) << kCountBits);
                const int k2 = j < n ? (int)((ev[j] >> kCountBits) >> level) : kInf;
                return max(k1, k2);
            }
            first = mid + 1;
            len = len - half - 1;
        }
    }
    return kInf;
}

// MINB = CTAs per SM the register allocation must allow: 2 (<= 64 registers) for whole-slice CTAs, 3
// (<= 42) for half-slice CTAs, whose shared memory lets three of them share an SM
template <int DPL, bool COMPAT, int MINB>
__global__ void __launch_bounds__(kMwWarps * 32, MINB) k_multitau_warp(MtArgs a, MwArgs m)
{
    constexpr int LG = DPL == 8 ? 3 : 2;
    constexpr int LO = DPL + 1;
    extern __shared__ __align__(16) uint32_t mw_smem[];
    // a CTA owns nrows = 32 / parts consecutive rows of slice s (fewer rows: a smaller stage, more CTAs
    // per SM); a warp-wide access then covers jstep = parts consecutive event ranks / delays
    const int s = blockIdx.x / m.parts;
    const int nrows = kSlice / m.parts, rbase = (blockIdx.x % m.parts) * nrows, jstep = m.parts;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rl = lane % nrows, jsub = lane / nrows;
    const int nwarps = blockDim.x >> 5;
    const int len = a.slice_len[s];
    if (len > m.len_cap) {  // CTA-uniform
        if (tid == 0) m.fallback[s] = 1;
        return;
    }
    uint32_t *evT = mw_smem;                              // [nrows][pitch_e]
    uint32_t *outS = evT + nrows * m.pitch_e;             // [3][nrows][pitch_t]
    uint32_t *wsm = outS + 3 * nrows * m.pitch_t + warp * m.warp_words;
    uint32_t *bw = wsm;                                   // [2 * len_cap + 64] words = 16-bit bins
    unsigned short *bh = reinterpret_cast<unsigned short *>(bw);
    uint32_t *tb = bw + 2 * m.len_cap + 64;
    uint32_t *tot = tb, *flim = tb + 32, *cnts = tb + 64, *cntml = tb + 96, *nlive = tb + 128, *sminS = tb + 160;

    const int F = a.sched.frames;
    const int nl = a.sched.n_levels;
    const int T = m.T;
    const int cnt0 = a.sched.count[0];
    const int lo0 = a.sched.lo[0];
    const int cumlast = m.lastl >= 1 ? DPL * (m.lastl - 1) + m.cnt_last : 0;

    // ---- phase 0: slice tile -> row-major shared memory (lane = row on the way in)
    const int my_len = a.row_len[s * kSlice + rbase + rl];
    {
        const uint32_t *g = reinterpret_cast<const uint32_t *>(a.store) + a.slice_base[s] + rbase + rl;
        uint32_t *dst = evT + rl * m.pitch_e;
#pragma unroll 4
        for (int j = warp * jstep + jsub; j < len; j += nwarps * jstep)
            if (j < my_len) dst[j] = g[(int64_t)j * kSlice];
    }
    __syncthreads();

    for (int rr = warp; rr < nrows; rr += nwarps) {
        const int n = __shfl_sync(kFull, my_len, rr);
        const uint32_t *ev = evT + rr * m.pitch_e;
        uint32_t *H = outS + rr * m.pitch_t;              // G2 numerators, later the G2 floats
        uint32_t *oIP = H + nrows * m.pitch_t;
        uint32_t *oIF = oIP + nrows * m.pitch_t;

        // ---- phase 1: count total, merge-level histogram, dead-level and IP / IF threshold histograms.
        // The IF thresholds t' << l ascend with the delay index and the IP thresholds (L_l - t') << l
        // descend, so an event only has to know the first delay it counts for: oIF[a] += c with
        // a = #{ti : t' << l <= f}, oIP[b] += c with b = #{ti : (L_l - t') << l > f} (closed forms from
        // the bit length of f and of F - f); phase 5 turns both into running sums.
        for (int t = lane; t < T; t += 32) {
            H[t] = 0u;
            oIP[t] = 0u;
            oIF[t] = 0u;
        }
        if (COMPAT) cntml[lane] = 0u;
        tot[lane] = 0u;  // until phase 2: counts of the events that die at level lane
        __syncwarp();
        uint32_t csum = 0;
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int i = c0 + lane;
            const uint32_t w = i < n ? ev[i] : 0u;
            const uint32_t c = w & ((1u << kCountBits) - 1u);
            csum += c;
            if (i < n) {
                const int f = (int)(w >> kCountBits);
                int ai;
                if (f < 2 * DPL) ai = min(f, cnt0);
                else {
                    const int ls = (32 - __clz(f)) - (LG + 1);
                    ai = cnt0 + DPL * (ls - 1) + (f >> ls) - DPL;
                }
                if (ai < T) atomicAdd(&oIF[ai], c);
                const int g = F - f;
                const int lc = (32 - __clz(g)) - (LG + 1);
                const int lf = lc - 2;  // levels 1..lf are full
                int bi = min(g - 1, cnt0) + (lf <= 0 ? 0 : (lf < m.lastl ? DPL * lf : cumlast));
#pragma unroll
                for (int q = 1; q >= 0; q--) {
                    const int l = lc - q;
                    if (l >= 1) {
                        const int cl = l < m.lastl ? DPL : (l == m.lastl ? m.cnt_last : 0);
                        const int u = (F >> l) - 1 - (f >> l);
                        bi += min(max(u - DPL, 0), cl);
                    }
                }
                if (bi < T) atomicAdd(&oIP[bi], c);
                // f >= lim_l  <=>  f >> l == F >> l  <=>  bitlength(f ^ F) <= l
                atomicAdd(&tot[32 - __clz(f ^ F)], c);
            }
            if (COMPAT && i >= 1 && i < n) {
                const int ml = 32 - __clz((int)((w ^ ev[i - 1]) >> kCountBits));
                atomicAdd(&cntml[ml], 1u);
            }
        }
        const uint32_t carry = __reduce_add_sync(kFull, csum);
        if (carry >= 65536u) {  // 32-bit numerators could overflow: the lane-per-row kernel redoes the slice
            if (lane == 0) m.fallback[s] = 1;
            continue;
        }
        __syncwarp();

        // first dense level: L_l <= ld_factor * n (levels below it cost one pair per partner, levels from it on one
        // bin per 2^l frames: the crossover lies where about every second bin is occupied)
        int ld;
        {
            const unsigned mk = __ballot_sync(kFull, lane >= 1 && lane < nl && (F >> lane) <= m.ld_factor * max(n, 1));
            ld = mk ? (__ffs(mk) - 1) : nl;
        }

        // ---- phase 2: per-level tables (lane = level)
        {
            const int l = lane;
            const bool lv_ok = l < nl;
            const int Ll = lv_ok ? (F >> l) : 0;
            const int liml = lv_ok ? (Ll << l) : 0;
            const int cnt_l = lv_ok ? a.sched.count[l] : 0;
            {  // PS(lim_l) = all counts - counts of the events dead at level l
                uint32_t dead = tot[l];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(kFull, dead, o);
                    if (lane >= o) dead += y;
                }
                __syncwarp();
                tot[l] = carry - dead;
            }
            cnts[l] = (l < ld) ? (uint32_t)cnt_l : 0u;
            uint32_t fl = (l < ld) ? (uint32_t)liml : 0u;
            if (COMPAT) {
                uint32_t ab = cntml[l];  // events that are not the first of their bin from level ml on
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(kFull, ab, o);
                    if (lane >= o) ab += y;
                }
                const int dropped = (n > 0 && lv_ok && (int)((ev[n - 1] >> kCountBits) >> l) >= Ll) ? 1 : 0;
                const int nv = n - (int)ab - dropped;
                int prev = __shfl_up_sync(kFull, nv, 1);
                if (lane == 0) prev = n;
                int sb = kInf;
                if (l >= 1 && lv_ok && nv < prev) sb = (int)((ev[nv] >> kCountBits) >> (l - 1));
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(kFull, sb, o);
                    if (lane >= o) sb = min(sb, y);
                }
                nlive[l] = (uint32_t)nv;
                sminS[l] = (uint32_t)sb;
                __syncwarp();
                const bool fire = l >= 1 && l < ld && cnt_l > 0 && nv < n && sb < Ll;
                unsigned mk = __ballot_sync(kFull, fire);
                while (mk) {
                    const int lv = __ffs(mk) - 1;
                    mk &= mk - 1;
                    const int ks = mw_stale_threshold(ev, n, nlive, lv, lane);
                    if (lane == lv && ks != kInf) fl = min(fl, (uint32_t)ks << lv);
                }
            }
            flim[l] = fl;
        }
        __syncwarp();

        // ---- phase 3: sparse levels, one pass over the event pairs i < j with f_j - f_i < dmax.
        // The pairs are flattened so that every lane carries one: pair p belongs to the event ("owner") with
        // Qc[k] <= p < Qc[k+1]; the owners that start inside a 32-pair chunk mark their first lane in a
        // bit mask, a lane's owner is the number of marks at or below it.
        {
            const uint32_t dmax = (uint32_t)(2 * DPL + 1) << (ld - 1);
            const uint32_t top0 = (uint32_t)(lo0 + cnt0 - 1);
            const uint32_t flim0 = flim[0];
            // events that have partners, compacted: Ic[k] = event index, Qc[k] = index of its first pair
            uint32_t *Qc = bw;
            uint32_t *Ic = bw + m.len_cap + 34;
            uint32_t npairs = 0;
            int nact = 0;
            const int topstep = n > 0 ? (1 << (31 - __clz(n))) : 1;
            for (int c0 = 0; c0 < n; c0 += 32) {
                const int i = c0 + lane;
                // pos = last index with f < f_i + dmax: galloping start (most events have < 8 partners)
                const uint32_t key = i < n ? (min((ev[i] >> kCountBits) + dmax, (uint32_t)F) << kCountBits) : 0u;
                int pos = i;
                if (__any_sync(kFull, i + 8 < n && ev[min(i + 8, n - 1)] < key)) {
                    for (int step = topstep; step >= 8; step >>= 1) {
                        const int k2 = pos + step;
                        if (k2 < n && ev[k2] < key) pos = k2;
                    }
                }
#pragma unroll
                for (int step = 4; step >= 1; step >>= 1) {
                    const int k2 = pos + step;
                    if (k2 < n && ev[k2] < key) pos = k2;
                }
                const uint32_t mi = i < n ? (uint32_t)(pos - i) : 0u;
                uint32_t x = mi;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(kFull, x, o);
                    if (lane >= o) x += y;
                }
                const unsigned act = __ballot_sync(kFull, mi > 0u);
                if (mi > 0u) {
                    const int k = nact + __popc(act & ((1u << lane) - 1u));
                    Qc[k] = npairs + x - mi;
                    Ic[k] = (uint32_t)i;
                }
                nact += __popc(act);
                npairs += __shfl_sync(kFull, x, 31);
            }
            Qc[nact + lane] = 0xffffffffu;  // sentinels: no further owner starts
            if (lane < 2) Qc[nact + 32 + lane] = 0xffffffffu;
            __syncwarp();
            int kbase = 0;  // compacted index of the owner of pair p0
            for (uint32_t p0 = 0; p0 < npairs; p0 += 32) {
                // owners that start inside this chunk set a bit at their first lane
                const uint32_t sl = Qc[kbase + 1 + lane] - p0;
                const unsigned bits = __reduce_or_sync(kFull, sl < 32u ? (1u << sl) : 0u);
                const int k = kbase + __popc(bits & (0xffffffffu >> (31 - lane)));
                kbase += __popc(bits);
                const uint32_t p = p0 + lane;
                if (p < npairs) {
                    const int i = (int)Ic[k];
                    const int j = i + 1 + (int)(p - Qc[k]);
                    const uint32_t wi = ev[i], wj = ev[j];
                    const uint32_t fi = wi >> kCountBits, fj = wj >> kCountBits;
                    const uint32_t d = fj - fi;
                    const uint32_t cc = (wi & ((1u << kCountBits) - 1u)) * (wj & ((1u << kCountBits) - 1u));
                    if (d <= top0 && d >= (uint32_t)lo0 && fj < flim0) atomicAdd(&H[d - lo0], cc);
                    if (d >= 2u * DPL) {
                        const int l0 = (32 - __clz((int)d)) - (LG + 1);  // d >> l0 in [dpl, 2 dpl)
                        const uint32_t b0 = (fj >> l0) - (fi >> l0) - LO;
                        if (b0 < cnts[l0] && fj < flim[l0]) atomicAdd(&H[cnt0 + (l0 - 1) * DPL + b0], cc);
                        const int l1 = l0 - 1;
                        if (l1 >= 1) {
                            const uint32_t b1 = (fj >> l1) - (fi >> l1);
                            if (b1 == 2u * DPL && (uint32_t)(DPL - 1) < cnts[l1] && fj < flim[l1])
                                atomicAdd(&H[cnt0 + (l1 - 1) * DPL + (DPL - 1)], cc);
                        }
                    }
                }
            }
            __syncwarp();
        }

        // ---- phase 4: dense levels, 16-bit bin arrays
        if (ld < nl) {
            int smin_run = COMPAT ? (int)sminS[ld] : kInf;
            for (int l = ld; l < nl; l++) {
                const int cnt_l = a.sched.count[l];
                if (cnt_l == 0) break;
                const int Ll = F >> l;
                if (l == ld) {
                    const int words = (Ll + 64) >> 1;
                    for (int t = lane; t < words; t += 32) bw[t] = 0u;
                    __syncwarp();
                    const uint32_t liml = (uint32_t)Ll << l;
                    for (int c0 = 0; c0 < n; c0 += 32) {
                        const int i = c0 + lane;
                        if (i < n) {
                            const uint32_t w = ev[i];
                            const uint32_t f = w >> kCountBits;
                            if (f < liml) {
                                const uint32_t key = f >> l;
                                atomicAdd(&bw[key >> 1], (w & ((1u << kCountBits) - 1u)) << ((key & 1u) * 16u));
                            }
                        }
                    }
                    __syncwarp();
                } else {
                    // B_l[t] = B_{l-1}[2t] + B_{l-1}[2t+1]; the key of the occupied level-(l-1) bin of rank
                    // n_l is the first stale slot this level leaves behind (compat)
                    const int want = COMPAT ? (int)nlive[l] : 0;
                    // (a level whose bins are all occupied cannot leave a stale slot below L_l: its key is >= n_l)
                    const bool track = COMPAT && nlive[l] < nlive[l - 1] && (int)nlive[l] < Ll;
                    int base = 0, cand = kInf;
                    for (int t0 = 0; t0 < Ll + 48; t0 += 32) {
                        const int t = t0 + lane;
                        const uint32_t word = t < Ll ? bw[t] : 0u;
                        const uint32_t vlo = word & 0xffffu, vhi = word >> 16;
                        if (track) {
                            const unsigned mlo = __ballot_sync(kFull, vlo != 0u), mhi = __ballot_sync(kFull, vhi != 0u);
                            const int c = __popc(mlo) + __popc(mhi);
                            if (want >= base && want < base + c) {
                                const unsigned lt = (1u << lane) - 1u;
                                const int rk = base + __popc(mlo & lt) + __popc(mhi & lt);
                                int mine = kInf;
                                if (vlo != 0u && rk == want) mine = 2 * t;
                                if (vhi != 0u && rk + (vlo != 0u ? 1 : 0) == want) mine = 2 * t + 1;
                                cand = min(cand, __reduce_min_sync(kFull, mine));
                            }
                            base += c;
                        }
                        __syncwarp();
                        bh[t] = (unsigned short)(vlo + vhi);
                        __syncwarp();
                    }
                    if (track) smin_run = min(smin_run, cand);
                }
                int klim = Ll;
                if (COMPAT && (int)nlive[l] < n && smin_run < Ll) klim = min(Ll, mw_stale_threshold(ev, n, nlive, l, lane));
                uint32_t acc[DPL];
#pragma unroll
                for (int k = 0; k < DPL; k++) acc[k] = 0u;
                if (klim == Ll) {
                    for (int t0 = 0; t0 < Ll - LO; t0 += 32) {
                        const int t = t0 + lane;
                        const uint32_t x = bh[t];
#pragma unroll
                        for (int k = 0; k < DPL; k++) acc[k] += x * (uint32_t)bh[t + LO + k];
                    }
                } else {
                    for (int t0 = 0; t0 < klim - LO; t0 += 32) {
                        const int t = t0 + lane;
                        const uint32_t x = bh[t];
#pragma unroll
                        for (int k = 0; k < DPL; k++)
                            if (t + LO + k < klim) acc[k] += x * (uint32_t)bh[t + LO + k];
                    }
                }
                uint32_t mine = 0u;
#pragma unroll
                for (int k = 0; k < DPL; k++) {
                    const uint32_t sum = __reduce_add_sync(kFull, acc[k]);
                    if (lane == k) mine = sum;
                }
                if (lane < cnt_l) H[cnt0 + (l - 1) * DPL + lane] = mine;
            }
        }
        __syncwarp();

        // ---- phase 5: running sums of the threshold histograms, one IEEE division per output (lane = delay)
        {
            uint32_t run = 0;  // both running sums in one word: each stays below 2^16
            for (int t0 = 0; t0 < T; t0 += 32) {
                const int ti = t0 + lane;
                uint32_t x = ti < T ? (oIF[ti] | (oIP[ti] << 16)) : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(kFull, x, o);
                    if (lane >= o) x += y;
                }
                x += run;
                run = __shfl_sync(kFull, x, 31);
                const uint32_t xc = x & 0xffffu, xd = x >> 16;
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
                    const float s2 = mw_pow2_neg(2 * l), s1 = mw_pow2_neg(l);
                    const uint32_t num = H[ti];
                    const uint32_t ipn = carry - xd;   // PS((L_l - t') << l)
                    const uint32_t ifn = tot[l] - xc;  // PS(lim_l) - PS(t' << l)
                    H[ti] = __float_as_uint(mw_scaled_div((float)num * s2, neff));
                    oIP[ti] = __float_as_uint(mw_scaled_div((float)ipn * s1, neff));
                    oIF[ti] = __float_as_uint(mw_scaled_div((float)ifn * s1, neff));
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- results: [3][nrows][T] stage -> [T][R_pad], 128 / parts contiguous bytes per (array, delay)
    {
        float *dst[3] = {a.G2, a.IP, a.IF};
        const int64_t r0 = (int64_t)s * kSlice + rbase + rl;
#pragma unroll
        for (int arr = 0; arr < 3; arr++) {
            const uint32_t *src = outS + arr * nrows * m.pitch_t + rl * m.pitch_t;
            float *d = dst[arr] + r0;
            for (int t = warp * jstep + jsub; t < T; t += nwarps * jstep) d[(int64_t)t * a.R_pad] = __uint_as_float(src[t]);
        }
    }
}

// The warp kernel covers: integer counts, dpl 4 or 8, frames below 2^20, and the regular
// schedule (level 0: delays 1.., level l >= 1: dpl+1..2dpl, no gaps).
bool multitau_warp_eligible(const xpcs_handle_s *h)
{
    if (h->kind != kPacked) return false;
    const int dpl = h->prm.delays_per_level;
    if (dpl != 4 && dpl != 8) return false;
    const Sched &sc = h->sched;
    if (sc.frames >= (1 << (32 - kCountBits)) || sc.n_levels > 24 || sc.n_levels < 1) return false;
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

template <int DPL, bool COMPAT, int MINB>
static int run_warp_b(xpcs_handle_s *h, MtArgs &a, MwArgs &m, size_t bytes, int warps)
{
    int rc = check_cuda(h, cudaFuncSetAttribute(k_multitau_warp<DPL, COMPAT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)bytes), "multitau_warp smem attr");
    if (rc) return rc;
    LaunchScope ls(h, "k_multitau_warp");
    k_multitau_warp<DPL, COMPAT, MINB><<<h->n_slices * m.parts, warps * 32, bytes, h->stream>>>(a, m);
    return XPCS_OK;
}

template <int DPL, bool COMPAT>
static int run_warp(xpcs_handle_s *h, MtArgs &a, MwArgs &m, size_t bytes, int warps)
{
    return m.parts >= 2 ? run_warp_b<DPL, COMPAT, 3>(h, a, m, bytes, warps) : run_warp_b<DPL, COMPAT, 2>(h, a, m, bytes, warps);
}

int launch_multitau_warp(xpcs_handle_s *h, MtArgs &a)
{
    int rc = ensure(h, h->d_mt_fallback, (size_t)(h->n_slices > 0 ? h->n_slices : 1), "multitau fallback flags");
    if (rc) return rc;
    cudaMemsetAsync(h->d_mt_fallback.p, 0, (size_t)(h->n_slices > 0 ? h->n_slices : 1), h->stream);
    if (h->n_slices == 0) return XPCS_OK;
    int smem_cap = 0;
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    MwArgs m{};
    m.fallback = h->d_mt_fallback.p;
    m.T = h->T;
    m.pitch_t = h->T | 1;
    for (int l = 1; l < h->sched.n_levels; l++)
        if (h->sched.count[l] > 0) {
            m.lastl = l;
            m.cnt_last = h->sched.count[l];
        }
    // bytes(len, warps, parts) = 4 * ((3 pitch_t + (len | 1)) * 32 / parts + warps * (2 len + 64 + tables)).
    // Occupancy decides (measured on C3: 16 / 24 / 32 / 48 warps per SM -> 6.3 / 5.1 / 4.2 / 4.1 ms; the
    // last step pays for its 42-register budget with ~10 % more instructions): three CTAs of
    // 16 warps per SM when half a slice per CTA makes them fit, else two CTAs of a whole slice -- with 16
    // warps each, or with 12 or 8 (a long delay schedule makes the result stage large) -- else one CTA
    // of 16 warps; longer slices go to the lane-per-row kernel
    auto bytes_for = [&](int len, int warps, int parts = 1) {
        return 4 * (((size_t)3 * m.pitch_t + (len | 1)) * (32 / parts) + (size_t)warps * (2 * (size_t)len + 64 + kMwTables));
    };
    const size_t budget2 = (size_t)(smem_cap + 1024) / 2 - 1024 - 512;  // two resident CTAs (1 KB reserved each)
    const size_t budget3 = (size_t)(smem_cap + 1024) / 3 - 1024 - 512;  // three
    int len_cap = h->max_row > 0 ? h->max_row : 1;
    {   // a few outlier rows (hot pixels) must not dictate the shared-memory budget of every CTA: slices more than
        // four times longer than the mean slice go to the lane-per-row kernel
        const int64_t mean_len = h->n_slices > 0 ? h->store_words / kSlice / h->n_slices : 0;
        len_cap = (int)std::min<int64_t>(len_cap, std::max<int64_t>(256, 4 * mean_len));
    }
    int warps = kMwWarps;
    m.parts = 1;
    if (bytes_for(len_cap, 16, 2) <= budget3) m.parts = 2;
    else if (bytes_for(len_cap, 16) > budget2) {
        if (bytes_for(len_cap, 12) <= budget2) warps = 12;
        else if (bytes_for(len_cap, 8) <= budget2) warps = 8;
        else {
            // long rows (C5 at >= 0.1 %: ~10^3 events): fewer rows per CTA.  Pick the (CTAs per slice, warps) pair that
            // keeps most warps resident on an SM; rows beyond every choice go to the lane-per-row kernel.
            const size_t budget1 = (size_t)smem_cap - 512;
            int best_rw = 0, best_parts = 1, best_warps = 16;
            for (int parts = 1; parts <= 32; parts *= 2)
                for (int w : {16, 12, 8, 4, 2, 1}) {
                    if (w > 32 / parts) continue;
                    const size_t b = bytes_for(len_cap, w, parts);
                    if (b > budget1) continue;
                    int ctas = (int)std::min<size_t>(32, ((size_t)smem_cap + 1024) / (b + 1024));
                    ctas = std::min(ctas, 48 / w > 0 ? 48 / w : 1);  // 42-register build: 1536 threads per SM
                    if (ctas * w > best_rw) {
                        best_rw = ctas * w;
                        best_parts = parts;
                        best_warps = w;
                    }
                }
            if (best_rw > 0) {
                m.parts = best_parts;
                warps = best_warps;
            } else
                while (len_cap > 1 && bytes_for(len_cap, 16) > budget1) len_cap = len_cap * 3 / 4;
        }
    }
    if (const char *e = getenv("XPCS_MW_WARPS")) {  // diagnostics: warps per CTA (4..16)
        const int w = atoi(e);
        if (w >= 4 && w <= kMwWarps) warps = w;
    }
    if (const char *e = getenv("XPCS_MW_PARTS")) {  // diagnostics: CTAs per slice (1, 2, 4)
        const int q = atoi(e);
        if (q == 1 || q == 2 || q == 4 || q == 8 || q == 16 || q == 32) m.parts = q;
    }
    if (bytes_for(len_cap, warps, m.parts) > (size_t)smem_cap) {  // T too large for the stage: everything falls back
        cudaMemsetAsync(h->d_mt_fallback.p, 1, (size_t)h->n_slices, h->stream);
        return XPCS_OK;
    }
    m.ld_factor = 4;
    if (const char *e = getenv("XPCS_MW_LD")) {  // diagnostics: 1..4
        const int q = atoi(e);
        if (q >= 1 && q <= 4) m.ld_factor = q;
    }
    m.len_cap = len_cap;
    m.pitch_e = len_cap | 1;
    m.warp_words = 2 * len_cap + 64 + kMwTables;
    const size_t bytes = bytes_for(len_cap, warps, m.parts);
    const bool compat = a.compat != 0;
    const int dpl = h->prm.delays_per_level;
    if (dpl == 8) rc = compat ? run_warp<8, true>(h, a, m, bytes, warps) : run_warp<8, false>(h, a, m, bytes, warps);
    else rc = compat ? run_warp<4, true>(h, a, m, bytes, warps) : run_warp<4, false>(h, a, m, bytes, warps);
    if (rc) return rc;
    return check_cuda(h, cudaGetLastError(), "k_multitau_warp");
}

}  // namespace xpcs
