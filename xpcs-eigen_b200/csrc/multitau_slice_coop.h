// multitau_slice_coop.h -- device-only front end of multitau_slice_core.h, shared by k_multitau_slice
// (multitau_slice.cu: packed integer rows) and k_multitau_slicef (multitau_slicef.cu: float rows, XS_NS = slf,
// XS_CB = 0).  The per-row routines are compiled for the device only here, and the rare stale-tail walk (one row in
// ~70 needs it on the bench workload, but then for thousands of instructions in a single lane) is given to the
// whole warp.
#pragma once
#include <stdint.h>

#ifndef XS_NS
#define XS_NS sl
#endif

namespace xpcs {
namespace XS_NS {
__device__ int coop_stale_walk(bool fire, const uint32_t *ev, int n, const uint32_t *nlive, int level);
}
}
#define XS_HD __device__ __forceinline__
#define XS_HD_CALL __device__ __noinline__
#define XS_WALK(fire, ev, n, nlive, level) coop_stale_walk((fire), (ev), (n), (nlive), (level))
#include "multitau_slice_core.h"

namespace xpcs {
namespace XS_NS {

// ---- the stale-tail walk of one row by a whole warp: lane_select_head / lane_key_at / lane_stale_threshold of
// multitau_slice_core.h with the lanes over the events (col = the row's column of the tile, stride 32 words)
__device__ __forceinline__ int coop_select_head(const uint32_t *col, int n, int level, int p, int lane)
{
    int base = 0;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int i = c0 + lane;
        bool head = false;
        if (i < n) head = (i == 0) || ((((col[i * kS] ^ col[(i - 1) * kS]) >> kCB) >> level) != 0u);
        const unsigned mk = __ballot_sync(0xffffffffu, head);
        const int c = __popc(mk);
        if (p < base + c) return c0 + (int)__fns(mk, 0, p - base + 1);
        base += c;
    }
    return n - 1;
}

__device__ __forceinline__ int coop_key_at(const uint32_t *col, int n, const uint32_t *nlive, int level, int p, int lane)
{
    int lv = level;
    if (p >= (int)nlive[level * kS]) {
        lv = level - 1;
        while (lv > 0 && (int)nlive[lv * kS] <= p) lv--;
    }
    const int i = coop_select_head(col, n, lv, p, lane);
    return (int)((col[i * kS] >> kCB) >> lv);
}

__device__ __noinline__ int coop_stale_threshold(const uint32_t *col, int n, const uint32_t *nlive, int level, int lane)
{
    const int nl = (int)nlive[level * kS];
    int first = 0, len = n;
    int curmin = kInfKey;
    while (len > 0) {
        const int half = len >> 1;
        const int mid = first + half;
        if (mid >= nl) {
            curmin = min(curmin, coop_key_at(col, n, nlive, level, mid, lane));
            len = half;
        } else {
            if (curmin != kInfKey && coop_key_at(col, n, nlive, level, mid, lane) > curmin) {
                const int k1 = coop_key_at(col, n, nlive, level, first, lane);
                const int j = lane_lower_bound(col, n, ((uint32_t)(curmin + 1) << level) << kCB);
                const int k2 = j < n ? (int)((col[j * kS] >> kCB) >> level) : kInfKey;
                return max(k1, k2);
            }
            first = mid + 1;
            len = len - half - 1;
        }
    }
    return kInfKey;
}

// K* of every lane that asks for it (fire), one row after the other with all 32 lanes on it
__device__ int coop_stale_walk(bool fire, const uint32_t *ev, int n, const uint32_t *nlive, int level)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    unsigned mk = __ballot_sync(0xffffffffu, fire);
    int ks = kInfKey;
    while (mk) {
        const int r = __ffs(mk) - 1;
        mk &= mk - 1;
        const int nr = __shfl_sync(0xffffffffu, n, r);
        const int k = coop_stale_threshold(ev - lane + r, nr, nlive - lane + r, level, lane);
        if (lane == r) ks = k;
    }
    return ks;
}

}  // namespace XS_NS
}  // namespace xpcs
