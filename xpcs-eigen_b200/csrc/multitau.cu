// multitau.cu -- per-pixel multi-tau correlator on the slice store.
//
// Replaces Corr::multiTau2 (reference corr.cpp:315-431).  One warp owns one slice of 32
// pixel rows, one lane one row; the slice is staged in shared memory column-interleaved
// (word j of lane L at [j*32 + L]: bank == lane, so every access is conflict free no matter
// how far the lanes have drifted apart), and every level is produced from that one load:
//   level l:  L_l = F >> l frames; keys halve, equal keys merge, keys >= L_l drop
//             (the in-place compaction of corr.cpp:349-390);
//             G2(tau') = sum over bin pairs at key distance tau'   (corr.cpp:397-411)
//             IP(tau') = total - sum_{key >= L-tau'}               (corr.cpp:403)
//             IF(tau') = total - sum_{key <  tau'}                 (corr.cpp:414-416)
//             each divided once, IEEE, by (L_l - tau')             (corr.cpp:420-424)
// Integer path (kPacked): counts stay integers, level-l values are count / 2^l, so the
// three numerators are exact and the single fp32 division reproduces the reference bit for
// bit.  Float path (kFloat): the reference's own fp32 operation order for
// the level merge, the G2 products and IP; IF via an fp64 total minus head.
// XPCS_COMPAT_STALE_TAIL: the compaction leaves the words beyond the live prefix exactly
// where the reference leaves its stale vector tail, and each candidate pair is confirmed by
// replaying std::lower_bound over that array.
#include "internal.h"

namespace xpcs {

__device__ __forceinline__ float pow2_neg(int e)  // 2^-e, exact
{
    return __int_as_float((127 - e) << 23);
}

__device__ __forceinline__ float scaled_div(float num, int neff)
{
    return neff > 0 ? __fdiv_rn(num, (float)neff) : num;
}

// ---- compat: the reference's binary search over the un-shrunk vector ------------------
// std::lower_bound (corr.cpp:406) runs over all n0 slots of the row: slots [0, n) hold the
// live level-l keys, slots [n, n0) whatever earlier levels left there.  For a target that
// sits at live position q the search is correct at every live probe; at a stale probe p >= n
// it must go left, which needs V[p] >= target.  So the pair is found iff every stale slot on
// the search path to q holds a value >= target, and the stale slots on that path depend on q
// only: walking the implicit search tree along the live/stale boundary gives a step function
// M(q) (minimum stale value probed before position q branches off), non-increasing in q.
// Keys increase with q, hence "lost" (M(q) < key[q]) is monotone: there is one threshold key
// K* such that exactly the targets with key >= K* are never found.
// KEYAT(p) decodes the value the reference sees in slot p.
template <typename KeyAt>
__device__ __forceinline__ int stale_tail_threshold(int n0, int n, KeyAt key_at)
{
    int first = 0, len = n0;
    int curmin = 0x7fffffff;
    while (len > 0) {
        const int half = len >> 1;
        const int mid = first + half;
        if (mid >= n) {  // stale probe: every live position of this range goes left of it
            curmin = min(curmin, key_at(mid));
            len = half;
        } else {         // live probe: positions [first, mid] leave the boundary path here
            if (curmin != 0x7fffffff && key_at(mid) > curmin) {
                int q = first;
                while (key_at(q) <= curmin) q++;
                return key_at(q);
            }
            first = mid + 1;
            len = len - half - 1;
        }
    }
    return 0x7fffffff;
}

// ---- integer path ------------------------------------------------------------------
// Dense-level pair sums.  When a level holds more than about one bin in six, walking the
// sparse list pair by pair costs far more than sliding a register window over ALL bins of
// the level: win[] holds the counts of bins t .. t+2*DPL (zero where the row has no event,
// zero from K* on in compat mode -- lost targets form a suffix, so sources from K* on have
// no surviving target either), and every bin t adds win[t] * win[t+DPL+1 .. t+2*DPL] to the
// DPL accumulators.  The window rotates by renaming (the loop is unrolled by its length).
template <int DPL>
__device__ __forceinline__ void dense_level_pairs(const uint32_t *col, int n, int cb, uint32_t cmask, int L,
                                                  int kstar, unsigned long long (&acc)[DPL])
{
    constexpr int W = 2 * DPL + 1;
    uint32_t win[W];
    int p = 0;
    uint32_t wp = n > 0 ? col[0] : 0u;
    int kp = n > 0 ? (int)(wp >> cb) : 0x7fffffff;
    auto fetch = [&](int key) -> uint32_t {  // keys are requested in ascending order, none skipped
        uint32_t v = 0;
        if (kp == key) {
            v = key < kstar ? (wp & cmask) : 0u;
            p++;
            if (p < n) {
                wp = col[p * kSlice];
                kp = (int)(wp >> cb);
            } else kp = 0x7fffffff;
        }
        return v;
    };
#pragma unroll
    for (int k = 0; k < W; k++) win[k] = fetch(k);
    for (int t0 = 0; t0 < L - DPL - 1; t0 += W) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            const uint32_t src = win[u];
#pragma unroll
            for (int d = 0; d < DPL; d++)
                acc[d] += (unsigned long long)src * (unsigned long long)win[(u + DPL + 1 + d) % W];
            win[u] = fetch(t0 + u + W);
        }
    }
}

template <bool COMPAT, int DPL>
__device__ __forceinline__ void row_multitau_packed(uint32_t *col, long long *acc, int n0, int r, const MtArgs &a)
{
    const int F = a.sched.frames;
    int n = n0;
    int L = F;
    int cb = kCountBits;
    int nhist[kMaxLevels];
    int stale_min = 0x7fffffff;
    for (int l = 0; l < a.sched.n_levels; l++) {
        if (l > 0) {
            // in-place compaction of corr.cpp:349-390: keys halve, equal keys merge, keys >= L drop
            L >>= 1;
            cb++;
            const uint32_t clr = ~(1u << (cb - 1));
            const uint32_t cmask = (1u << cb) - 1u;
            int m = 0;
            uint32_t prev = 0xffffffffu;
            uint32_t cur = 0;
            for (int j = 0; j < n; j++) {
                const uint32_t w = col[j * kSlice] & clr;
                const uint32_t key = w >> cb;
                if ((int)key >= L) break;
                if (key == prev) cur += (w & cmask);
                else {
                    if (m > 0) col[(m - 1) * kSlice] = cur;
                    cur = w;
                    m++;
                    prev = key;
                }
            }
            if (m > 0) col[(m - 1) * kSlice] = cur;
            if (COMPAT && m < n) stale_min = min(stale_min, (int)(col[m * kSlice] >> (cb - 1)));
            n = m;
        }
        if (COMPAT) nhist[l] = n;
        const int cnt = a.sched.count[l];
        if (cnt == 0) continue;
        const uint32_t cmask = (1u << cb) - 1u;
        const int lo = a.sched.lo[l];
        const int top = lo + cnt - 1;  // largest level-local delay needed
        const int first = a.sched.first[l];
        int kstar = 0x7fffffff;
        if (COMPAT && l > 0 && stale_min < L && n < n0) {
            const int level = l;
            kstar = stale_tail_threshold(n0, n, [&](int p) -> int {
                if (p < n) return (int)(col[p * kSlice] >> cb);
                int lv = level - 1;
                while (lv > 0 && nhist[lv] <= p) lv--;
                return (int)(col[p * kSlice] >> (kCountBits + lv));
            });
        }
        const float s2 = pow2_neg(2 * l), s1 = pow2_neg(l);
        // ---- G2: pairs of bins at key distance lo..top (the later bin is the search target)
        bool dense = false;
        if (DPL > 0 && l > 0 && lo == DPL + 1)
            dense = (long long)__reduce_add_sync(0xffffffffu, n) * 6 > (long long)L * 32;
        long long total = 0;
        if (DPL > 0 && dense) {
            unsigned long long pairs[DPL > 0 ? DPL : 1];
#pragma unroll
            for (int d = 0; d < DPL; d++) pairs[d] = 0ull;
            dense_level_pairs<(DPL > 0 ? DPL : 1)>(col, n, cb, cmask, L, kstar, pairs);
            for (int i = 0; i < n; i++) total += (long long)(col[i * kSlice] & cmask);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < DPL; k++)
                if (k < cnt)
                    a.G2[(int64_t)(first + k) * a.R_pad + r] = scaled_div((float)(long long)pairs[k] * s2, L - (lo + k));
        } else {
            for (int d = lo; d <= top; d++) acc[d * kSlice] = 0;
            uint32_t wnext = n > 0 ? col[0] : 0u;
            for (int i = 0; i < n; i++) {
                const uint32_t wi = wnext;
                const int ki = (int)(wi >> cb);
                const int ci = (int)(wi & cmask);
                total += ci;
                if (i + 1 < n) wnext = col[(i + 1) * kSlice];
                uint32_t wj = wnext;
                for (int j = i + 1; j < n;) {
                    const int kj = (int)(wj >> cb);
                    const int d = kj - ki;
                    if (d > top) break;
                    if (d >= lo && kj < kstar) acc[d * kSlice] += (long long)ci * (long long)(wj & cmask);
                    if (++j < n) wj = col[j * kSlice];
                }
            }
            __syncwarp();
            for (int k = 0; k < cnt; k++) {
                const int tp = lo + k;
                a.G2[(int64_t)(first + k) * a.R_pad + r] = scaled_div((float)acc[tp * kSlice] * s2, L - tp);
            }
        }
        // ---- IF: total minus the bins with key < tau'
        for (int d = 0; d < top; d++) acc[d * kSlice] = 0;
        for (int i = 0; i < n; i++) {
            const uint32_t wi = col[i * kSlice];
            const int ki = (int)(wi >> cb);
            if (ki >= top) break;
            acc[ki * kSlice] = (long long)(wi & cmask);
        }
        {
            long long run = 0;
            for (int tp = 1; tp <= top; tp++) {
                run += acc[(tp - 1) * kSlice];
                if (tp >= lo)
                    a.IF[(int64_t)(first + tp - lo) * a.R_pad + r] = scaled_div((float)(total - run) * s1, L - tp);
            }
        }
        // ---- IP: total minus the bins with key >= L - tau'
        for (int d = 0; d < top; d++) acc[d * kSlice] = 0;
        for (int i = n - 1; i >= 0; i--) {
            const uint32_t wi = col[i * kSlice];
            const int x = L - 1 - (int)(wi >> cb);
            if (x >= top) break;
            acc[x * kSlice] = (long long)(wi & cmask);
        }
        {
            long long run = 0;
            for (int tp = 1; tp <= top; tp++) {
                run += acc[(tp - 1) * kSlice];
                if (tp >= lo)
                    a.IP[(int64_t)(first + tp - lo) * a.R_pad + r] = scaled_div((float)(total - run) * s1, L - tp);
            }
        }
    }
}

// ---- float path --------------------------------------------------------------------
typedef unsigned long long u64;

__device__ __forceinline__ int f_key(u64 w) { return (int)(w >> 32); }
__device__ __forceinline__ float f_val(u64 w) { return __uint_as_float((uint32_t)w); }
__device__ __forceinline__ u64 f_make(int key, float v)
{
    return ((u64)(uint32_t)key << 32) | (u64)__float_as_uint(v);
}

// Register-window pair sums of a dense level, float rows.  Per lag the products are added in
// bin order, which is the reference's order (sources ascending, corr.cpp:397-411); the zero
// products of empty bins leave an fp32 sum unchanged, so the result is bit-identical to the
// sparse walk.
template <int DPL>
__device__ __forceinline__ void dense_level_pairs_float(const u64 *col, int n, int L, int kstar, float (&acc)[DPL])
{
    constexpr int W = 2 * DPL + 1;
    float win[W];
    int p = 0;
    u64 wp = n > 0 ? col[0] : 0ull;
    int kp = n > 0 ? f_key(wp) : 0x7fffffff;
    auto fetch = [&](int key) -> float {
        float v = 0.0f;
        if (kp == key) {
            v = key < kstar ? f_val(wp) : 0.0f;
            p++;
            if (p < n) {
                wp = col[p * kSlice];
                kp = f_key(wp);
            } else kp = 0x7fffffff;
        }
        return v;
    };
#pragma unroll
    for (int k = 0; k < W; k++) win[k] = fetch(k);
    for (int t0 = 0; t0 < L - DPL - 1; t0 += W) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            const float src = win[u];
#pragma unroll
            for (int d = 0; d < DPL; d++) acc[d] = __fadd_rn(acc[d], __fmul_rn(src, win[(u + DPL + 1 + d) % W]));
            win[u] = fetch(t0 + u + W);
        }
    }
}

template <bool COMPAT, int DPL>
__device__ void row_multitau_float(u64 *col, float *acc, int n0, int r, const MtArgs &a)
{
    const int F = a.sched.frames;
    const int hi = a.hi;
    int n = n0;
    int L = F;
    int stale_min = 0x7fffffff;
    for (int l = 0; l < a.sched.n_levels; l++) {
        if (l > 0) {
            L >>= 1;
            int m = 0;
            int prev = -1;
            for (int j = 0; j < n; j++) {
                const u64 w = col[j * kSlice];
                const int key = f_key(w) >> 1;
                if (key >= L) break;
                if (key == prev) {
                    const u64 q = col[(m - 1) * kSlice];
                    col[(m - 1) * kSlice] = f_make(key, __fadd_rn(f_val(q), f_val(w)));
                } else {
                    col[m * kSlice] = f_make(key, f_val(w));
                    m++;
                    prev = key;
                }
            }
            if (COMPAT && m < n) stale_min = min(stale_min, f_key(col[m * kSlice]));
            for (int j = 0; j < m; j++) {  // val /= 2.0f (corr.cpp:388-389), exact
                const u64 w = col[j * kSlice];
                col[j * kSlice] = f_make(f_key(w), __fmul_rn(f_val(w), 0.5f));
            }
            n = m;
        }
        const int cnt = a.sched.count[l];
        if (cnt == 0) continue;
        const int lo = a.sched.lo[l];
        const int top = lo + cnt - 1;
        const int first = a.sched.first[l];
        int kstar = 0x7fffffff;
        if (COMPAT && l > 0 && stale_min < L && n < n0)
            kstar = stale_tail_threshold(n0, n, [&](int p) -> int { return f_key(col[p * kSlice]); });
        // ---- G2 in the reference's order: for each source ascending, product then add
        double total = 0.0;
        bool dense = false;
        if (DPL > 0 && l > 0 && lo == DPL + 1)
            dense = (long long)__reduce_add_sync(0xffffffffu, n) * 6 > (long long)L * 32;
        if (DPL > 0 && dense) {
            float pairs[DPL > 0 ? DPL : 1];
#pragma unroll
            for (int d = 0; d < DPL; d++) pairs[d] = 0.0f;
            dense_level_pairs_float<(DPL > 0 ? DPL : 1)>(col, n, L, kstar, pairs);
            for (int i = 0; i < n; i++) total += (double)f_val(col[i * kSlice]);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < DPL; k++)
                if (k < cnt) a.G2[(int64_t)(first + k) * a.R_pad + r] = scaled_div(pairs[k], L - (lo + k));
        } else {
            for (int d = lo; d <= top; d++) acc[d * kSlice] = 0.0f;
            for (int i = 0; i < n; i++) {
                const u64 wi = col[i * kSlice];
                const int ki = f_key(wi);
                const float vi = f_val(wi);
                total += (double)vi;
                for (int j = i + 1; j < n; j++) {
                    const u64 wj = col[j * kSlice];
                    const int kj = f_key(wj);
                    const int d = kj - ki;
                    if (d > top) break;
                    if (d >= lo && kj < kstar) acc[d * kSlice] = __fadd_rn(acc[d * kSlice], __fmul_rn(vi, f_val(wj)));
                }
            }
            __syncwarp();
            for (int k = 0; k < cnt; k++) {
                const int tp = lo + k;
                a.G2[(int64_t)(first + k) * a.R_pad + r] = scaled_div(acc[tp * kSlice], L - tp);
            }
        }
        // ---- IP: the running fp32 prefix sum at the moment the key reaches L - tau'
        {
            float run = 0.0f;
            int tp = hi;
            for (int i = 0; i < n; i++) {
                const u64 wi = col[i * kSlice];
                const int ki = f_key(wi);
                while (tp >= 1 && ki >= L - tp) {
                    acc[tp * kSlice] = run;
                    tp--;
                }
                run = __fadd_rn(run, f_val(wi));
            }
            while (tp >= 1) {
                acc[tp * kSlice] = run;
                tp--;
            }
            for (int k = 0; k < cnt; k++) {
                const int t2 = lo + k;
                a.IP[(int64_t)(first + k) * a.R_pad + r] = scaled_div(acc[t2 * kSlice], L - t2);
            }
        }
        // ---- IF in the reference's order (corr.cpp:414-416): for every delay its own sequential fp32
        // chain over the bins with key >= tau' -- the chains differ in where they start, hence in
        // every rounding after that, and a long row's chain carries more than 1e-5 of rounding,
        // so only the same order reproduces the reference
        for (int tp = lo; tp <= top; tp++) acc[tp * kSlice] = 0.0f;
        for (int i = 0; i < n; i++) {
            const u64 wi = col[i * kSlice];
            const int ki = f_key(wi);
            const float vi = f_val(wi);
            const int last = min(ki, top);
            for (int tp = lo; tp <= last; tp++) acc[tp * kSlice] = __fadd_rn(acc[tp * kSlice], vi);
        }
        for (int k = 0; k < cnt; k++)
            a.IF[(int64_t)(first + k) * a.R_pad + r] = scaled_div(acc[(lo + k) * kSlice], L - (lo + k));
    }
}

template <int KIND, bool COMPAT, int DPL>
__global__ void __launch_bounds__(32) k_multitau(MtArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = blockIdx.x;
    if (a.only_flagged && !a.only_flagged[s]) return;  // slice already done by k_multitau_warp
    const int lane = threadIdx.x;
    const int r = s * kSlice + lane;
    const int len = a.slice_len[s];
    const int n0 = a.row_len[r];
    const bool in_smem = len <= a.smem_len;
    if (KIND == kPacked) {
        uint32_t *g = reinterpret_cast<uint32_t *>(a.store) + a.slice_base[s] + lane;
        long long *acc = reinterpret_cast<long long *>(smem_raw) + lane;
        if (in_smem) {  // two call sites so that the common one compiles to LDS/STS
            uint32_t *col = reinterpret_cast<uint32_t *>(smem_raw + (size_t)(a.hi + 1) * kSlice * 8) + lane;
            for (int j = 0; j < len; j++)
                if (j < n0) col[j * kSlice] = g[(int64_t)j * kSlice];
            row_multitau_packed<COMPAT, DPL>(col, acc, n0, r, a);
        } else row_multitau_packed<COMPAT, 0>(g, acc, n0, r, a);
    } else {
        u64 *g = reinterpret_cast<u64 *>(a.store) + a.slice_base[s] + lane;
        u64 *col = g;
        float *acc = reinterpret_cast<float *>(smem_raw) + lane;
        if (in_smem) {
            col = reinterpret_cast<u64 *>(smem_raw + (size_t)(a.hi + 1) * kSlice * 8) + lane;
            for (int j = 0; j < len; j++)
                if (j < n0) col[j * kSlice] = g[(int64_t)j * kSlice];
        }
        row_multitau_float<COMPAT, DPL>(col, acc, n0, r, a);
    }
}

// [T][R_pad] row-permuted -> [T][P] detector order (zeros elsewhere, pre-cleared by caller)
__global__ void k_unpermute(const float *__restrict__ src, float *__restrict__ dst,
                            const int *__restrict__ pixel_of_row, int R, int R_pad, int P)
{
    const int t = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    dst[(int64_t)t * P + pixel_of_row[r]] = src[(int64_t)t * R_pad + r];
}

template <int KIND, bool COMPAT, int DPL>
static int run_multitau(xpcs_handle_s *h, MtArgs &a)
{
    int smem_cap = 0;
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    smem_cap -= 1024;
    const size_t wbytes = KIND == kPacked ? 4 : 8;
    const size_t acc_bytes = (size_t)(a.hi + 1) * kSlice * 8;
    int smem_len = h->max_row;
    if (acc_bytes + (size_t)smem_len * kSlice * wbytes > (size_t)smem_cap)
        smem_len = (int)((smem_cap - acc_bytes) / (kSlice * wbytes));
    a.smem_len = smem_len;
    if (h->max_row > smem_len) h->rows_consumed = true;  // long rows are compacted in place
    const size_t bytes = acc_bytes + (size_t)smem_len * kSlice * wbytes;
    int rc = check_cuda(h, cudaFuncSetAttribute(k_multitau<KIND, COMPAT, DPL>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                        "multitau smem attr");
    if (rc) return rc;
    if (h->n_slices > 0) {
        LaunchScope ls(h, "k_multitau");
        k_multitau<KIND, COMPAT, DPL><<<h->n_slices, 32, bytes, h->stream>>>(a);
    }
    return check_cuda(h, cudaGetLastError(), "k_multitau");
}

int launch_multitau(xpcs_handle_s *h)
{
    int rc;
    const size_t n = (size_t)h->T * h->R_pad;
    if ((rc = ensure(h, h->d_G2, n, "G2"))) return rc;
    if ((rc = ensure(h, h->d_IP, n, "IP"))) return rc;
    if ((rc = ensure(h, h->d_IF, n, "IF"))) return rc;
    MtArgs a{};
    a.store = h->d_store.p;
    a.slice_base = h->d_slice_base.p;
    a.slice_len = h->d_slice_len.p;
    a.row_len = h->d_row_len.p;
    a.G2 = h->d_G2.p;
    a.IP = h->d_IP.p;
    a.IF = h->d_IF.p;
    a.R_pad = h->R_pad;
    a.n_slices = h->n_slices;
    a.hi = 2 * h->prm.delays_per_level;
    a.compat = (h->prm.compat_flags & XPCS_COMPAT_STALE_TAIL) ? 1 : 0;
    a.sched = h->sched;
    a.only_flagged = nullptr;
    h->mt_warp_ran = false;
    if (multitau_slice_eligible(h) && !(h->prm.compat_flags & XPCS_FLAG_LANE_MULTITAU)) {
        // short rows: lane = row, warps = tasks (multitau_slice.cu); it flags the slices it leaves like the
        // warp-per-row kernel does
        if ((rc = launch_multitau_slice(h, a))) return rc;
        a.only_flagged = h->d_mt_fallback.p;
        h->mt_warp_ran = true;
    } else if (multitau_warp_eligible(h) && !(h->prm.compat_flags & XPCS_FLAG_LANE_MULTITAU)) {
        // warp-per-row kernel first; it flags the slices it leaves (rows too long for its shared
        // memory, or counts too large for 32-bit numerators) and the lane-per-row kernel redoes those
        if ((rc = launch_multitau_warp(h, a))) return rc;
        a.only_flagged = h->d_mt_fallback.p;
        h->mt_warp_ran = true;
    }
    if (multitau_slicef_eligible(h) && !(h->prm.compat_flags & XPCS_FLAG_LANE_MULTITAU)) {
        // float rows whose slices fit a shared-memory tile: lane = row, warps = tasks (multitau_slicef.cu); the slices it
        // flags (hot pixels) go to the warp-per-row kernel, and what that one leaves to the lane-per-row kernel
        if ((rc = launch_multitau_slicef(h, a))) return rc;
        if ((rc = launch_multitau_warpf(h, a, true))) return rc;
        a.only_flagged = h->d_mt_fallback.p;
        h->mt_warp_ran = true;
    } else if (multitau_warpf_eligible(h) && !(h->prm.compat_flags & XPCS_FLAG_LANE_MULTITAU)) {
        // float rows: same split between the warp-per-row kernel and the lane-per-row one
        if ((rc = launch_multitau_warpf(h, a))) return rc;
        a.only_flagged = h->d_mt_fallback.p;
        h->mt_warp_ran = true;
    }
    if (h->kind == kPacked) {
        // the register-window path for dense levels is instantiated for the usual delays-per-level
        const int dpl = h->prm.delays_per_level;
        if (dpl == 8) return a.compat ? run_multitau<kPacked, true, 8>(h, a) : run_multitau<kPacked, false, 8>(h, a);
        if (dpl == 4) return a.compat ? run_multitau<kPacked, true, 4>(h, a) : run_multitau<kPacked, false, 4>(h, a);
        return a.compat ? run_multitau<kPacked, true, 0>(h, a) : run_multitau<kPacked, false, 0>(h, a);
    }
    {
        const int dpl = h->prm.delays_per_level;
        if (dpl == 8) return a.compat ? run_multitau<kFloat, true, 8>(h, a) : run_multitau<kFloat, false, 8>(h, a);
        if (dpl == 4) return a.compat ? run_multitau<kFloat, true, 4>(h, a) : run_multitau<kFloat, false, 4>(h, a);
    }
    return a.compat ? run_multitau<kFloat, true, 0>(h, a) : run_multitau<kFloat, false, 0>(h, a);
}

int launch_unpermute(xpcs_handle_s *h, const float *d_src, float *d_dst)
{
    cudaMemsetAsync(d_dst, 0, sizeof(float) * (size_t)h->T * h->P, h->stream);
    if (h->R > 0) {
        LaunchScope ls(h, "k_unpermute");
        dim3 grid((h->R + 255) / 256, h->T);
        k_unpermute<<<grid, 256, 0, h->stream>>>(d_src, d_dst, h->d_pixel_of_row.p, h->R, h->R_pad, h->P);
    }
    return check_cuda(h, cudaGetLastError(), "k_unpermute");
}

}  // namespace xpcs
