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

struct MtArgs {
    void *store;
    const int64_t *slice_base;
    const int *slice_len;
    const int *row_len;
    float *G2, *IP, *IF;
    int R_pad, n_slices, smem_len, hi, compat;
    Sched sched;
};

__device__ __forceinline__ float pow2_neg(int e)  // 2^-e, exact
{
    return __int_as_float((127 - e) << 23);
}

__device__ __forceinline__ float scaled_div(float num, int neff)
{
    return neff > 0 ? __fdiv_rn(num, (float)neff) : num;
}

// ---- integer path ------------------------------------------------------------------
// Replays the probe sequence of libstdc++'s std::lower_bound over the n0 slots of the row:
// slots below n hold level-l keys, slots at or above n hold what earlier levels left there.
__device__ __forceinline__ int packed_slot_key(const uint32_t *col, int p, int n, int cb, int level,
                                               const int *nhist)
{
    if (p < n) return (int)(col[p * kSlice] >> cb);
    int lv = level - 1;
    while (lv > 0 && nhist[lv] <= p) lv--;
    return (int)(col[p * kSlice] >> (kCountBits + lv));
}

__device__ bool packed_ref_search_hits(const uint32_t *col, int n0, int n, int cb, int level,
                                       const int *nhist, int target)
{
    int first = 0, len = n0;
    while (len > 0) {
        int half = len >> 1;
        int mid = first + half;
        if (packed_slot_key(col, mid, n, cb, level, nhist) < target) {
            first = mid + 1;
            len = len - half - 1;
        } else len = half;
    }
    return first < n && (int)(col[first * kSlice] >> cb) == target;
}

template <bool COMPAT>
__device__ void row_multitau_packed(uint32_t *col, long long *acc, int n0, int r, const MtArgs &a)
{
    const int F = a.sched.frames;
    const int hi = a.hi;
    int n = n0;
    int L = F;
    int cb = kCountBits;
    int nhist[kMaxLevels];
    int stale_min = 0x7fffffff;
    for (int l = 0; l < a.sched.n_levels; l++) {
        if (l > 0) {
            L >>= 1;
            cb++;
            const uint32_t clr = ~(1u << (cb - 1));
            const uint32_t cmask = (1u << cb) - 1u;
            int m = 0;
            uint32_t prev = 0xffffffffu;
            for (int j = 0; j < n; j++) {
                uint32_t w = col[j * kSlice] & clr;
                uint32_t key = w >> cb;
                if ((int)key >= L) break;
                if (key == prev) col[(m - 1) * kSlice] += (w & cmask);
                else {
                    col[m * kSlice] = w;
                    m++;
                    prev = key;
                }
            }
            if (COMPAT && m < n) stale_min = min(stale_min, (int)(col[m * kSlice] >> (cb - 1)));
            n = m;
        }
        if (COMPAT) nhist[l] = n;
        const int cnt = a.sched.count[l];
        if (cnt == 0) continue;
        const uint32_t cmask = (1u << cb) - 1u;
        const int lo = a.sched.lo[l];
        const int first = a.sched.first[l];
        const bool verify = COMPAT && l > 0 && stale_min < L;
        // ---- G2: pairs of bins at key distance <= hi
        for (int d = 0; d <= hi; d++) acc[d * kSlice] = 0;
        long long total = 0;
        for (int i = 0; i < n; i++) {
            const uint32_t wi = col[i * kSlice];
            const int ki = (int)(wi >> cb);
            const int ci = (int)(wi & cmask);
            total += ci;
            for (int j = i + 1; j < n; j++) {
                const uint32_t wj = col[j * kSlice];
                const int d = (int)(wj >> cb) - ki;
                if (d > hi) break;
                if (COMPAT && verify && d >= lo &&
                    !packed_ref_search_hits(col, n0, n, cb, l, nhist, ki + d))
                    continue;
                acc[d * kSlice] += (long long)ci * (long long)(wj & cmask);
            }
        }
        const float s2 = pow2_neg(2 * l), s1 = pow2_neg(l);
        for (int k = 0; k < cnt; k++) {
            const int tp = lo + k;
            a.G2[(int64_t)(first + k) * a.R_pad + r] = scaled_div((float)acc[tp * kSlice] * s2, L - tp);
        }
        // ---- IF: total minus the bins with key < tau'
        for (int d = 0; d <= hi; d++) acc[d * kSlice] = 0;
        for (int i = 0; i < n; i++) {
            const uint32_t wi = col[i * kSlice];
            const int ki = (int)(wi >> cb);
            if (ki >= hi) break;
            acc[ki * kSlice] += (int)(wi & cmask);
        }
        {
            long long run = 0;
            for (int tp = 1; tp < lo + cnt; tp++) {
                run += acc[(tp - 1) * kSlice];
                if (tp >= lo)
                    a.IF[(int64_t)(first + tp - lo) * a.R_pad + r] = scaled_div((float)(total - run) * s1, L - tp);
            }
        }
        // ---- IP: total minus the bins with key >= L - tau'
        for (int d = 0; d <= hi; d++) acc[d * kSlice] = 0;
        for (int i = n - 1; i >= 0; i--) {
            const uint32_t wi = col[i * kSlice];
            const int x = L - 1 - (int)(wi >> cb);
            if (x >= hi) break;
            acc[x * kSlice] += (int)(wi & cmask);
        }
        {
            long long run = 0;
            for (int tp = 1; tp < lo + cnt; tp++) {
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

__device__ bool float_ref_search_hits(const u64 *col, int n0, int n, int target)
{
    int first = 0, len = n0;
    while (len > 0) {
        int half = len >> 1;
        int mid = first + half;
        if (f_key(col[mid * kSlice]) < target) {
            first = mid + 1;
            len = len - half - 1;
        } else len = half;
    }
    return first < n && f_key(col[first * kSlice]) == target;
}

template <bool COMPAT>
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
        const int first = a.sched.first[l];
        const bool verify = COMPAT && l > 0 && stale_min < L;
        // ---- G2 in the reference's order: for each source ascending, product then add
        for (int d = 0; d <= hi; d++) acc[d * kSlice] = 0.0f;
        double total = 0.0;
        for (int i = 0; i < n; i++) {
            const u64 wi = col[i * kSlice];
            const int ki = f_key(wi);
            const float vi = f_val(wi);
            total += (double)vi;
            for (int j = i + 1; j < n; j++) {
                const u64 wj = col[j * kSlice];
                const int d = f_key(wj) - ki;
                if (d > hi) break;
                if (COMPAT && verify && d >= lo && !float_ref_search_hits(col, n0, n, ki + d)) continue;
                acc[d * kSlice] = __fadd_rn(acc[d * kSlice], __fmul_rn(vi, f_val(wj)));
            }
        }
        for (int k = 0; k < cnt; k++) {
            const int tp = lo + k;
            a.G2[(int64_t)(first + k) * a.R_pad + r] = scaled_div(acc[tp * kSlice], L - tp);
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
        // ---- IF: fp64 total minus the head bins (keys are distinct, one value per key)
        for (int d = 0; d <= hi; d++) acc[d * kSlice] = 0.0f;
        for (int i = 0; i < n; i++) {
            const u64 wi = col[i * kSlice];
            const int ki = f_key(wi);
            if (ki >= hi) break;
            acc[ki * kSlice] = f_val(wi);
        }
        {
            double run = 0.0;
            for (int tp = 1; tp < lo + cnt; tp++) {
                run += (double)acc[(tp - 1) * kSlice];
                if (tp >= lo)
                    a.IF[(int64_t)(first + tp - lo) * a.R_pad + r] = scaled_div((float)(total - run), L - tp);
            }
        }
    }
}

template <int KIND, bool COMPAT>
__global__ void __launch_bounds__(32) k_multitau(MtArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = blockIdx.x;
    const int lane = threadIdx.x;
    const int r = s * kSlice + lane;
    const int len = a.slice_len[s];
    const int n0 = a.row_len[r];
    const bool in_smem = len <= a.smem_len;
    if (KIND == kPacked) {
        uint32_t *g = reinterpret_cast<uint32_t *>(a.store) + a.slice_base[s] + lane;
        uint32_t *col = g;
        long long *acc = reinterpret_cast<long long *>(smem_raw) + lane;
        if (in_smem) {
            col = reinterpret_cast<uint32_t *>(smem_raw + (size_t)(a.hi + 1) * kSlice * 8) + lane;
            for (int j = 0; j < len; j++)
                if (j < n0) col[j * kSlice] = g[(int64_t)j * kSlice];
        }
        row_multitau_packed<COMPAT>(col, acc, n0, r, a);
    } else {
        u64 *g = reinterpret_cast<u64 *>(a.store) + a.slice_base[s] + lane;
        u64 *col = g;
        float *acc = reinterpret_cast<float *>(smem_raw) + lane;
        if (in_smem) {
            col = reinterpret_cast<u64 *>(smem_raw + (size_t)(a.hi + 1) * kSlice * 8) + lane;
            for (int j = 0; j < len; j++)
                if (j < n0) col[j * kSlice] = g[(int64_t)j * kSlice];
        }
        row_multitau_float<COMPAT>(col, acc, n0, r, a);
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

template <int KIND, bool COMPAT>
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
    int rc = check_cuda(h, cudaFuncSetAttribute(k_multitau<KIND, COMPAT>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                        "multitau smem attr");
    if (rc) return rc;
    if (h->n_slices > 0) {
        LaunchScope ls(h, "k_multitau");
        k_multitau<KIND, COMPAT><<<h->n_slices, 32, bytes, h->stream>>>(a);
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
    if (h->kind == kPacked)
        return a.compat ? run_multitau<kPacked, true>(h, a) : run_multitau<kPacked, false>(h, a);
    return a.compat ? run_multitau<kFloat, true>(h, a) : run_multitau<kFloat, false>(h, a);
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
