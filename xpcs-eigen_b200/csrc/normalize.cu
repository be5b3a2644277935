// normalize.cu -- q-bin normalisation: static-partition means, normalised g2 per dynamic
// bin and its standard error.  Replaces Corr::normalizeG2s (reference corr.cpp:927-1091).
//
// Rows are stored in (dq, sq, pixel) order, so every static segment is a contiguous row
// range of the [T][R_pad] correlator arrays.
//  k_segment_reduce  one CTA per (segment, chunk of 32 delays); tiles of 32 rows stream through
//                    a cp.async ring; warp 0 (lane = delay) folds them in row order, so the three
//                    fp32 sums run in exactly the reference's order (corr.cpp:966-991: sequential
//                    += over the pixels of a static bin); warps 1..3 accumulate, in fp64, the sum
//                    and sum of squares of the per-pixel x = G2/(IP*IF) (NaN -> 0, :1066-1070).
//  k_normalize_finish  per (dynamic bin, delay): NaN-skipping fp32 mean of the static-bin g2
//                    (corr.cpp:1021-1039) and stderr = sqrt(1/n)*sqrt(M2/n) (corr.cpp:1083-1086)
//                    with M2 = sum x^2 - (sum x)^2/n from the fp64 partials.  The reference runs a
//                    float32 Welford chain over all pixels of the bin; the fp64 sums agree with
//                    it to ~1e-6 relative and do not serialise 10^4..10^5 steps.
// Partials buffer (doubles): [nseg_total][T] segment g2, then [nseg_total][T] sum x, then
// [nseg_total][T] sum x^2.  Entries of segments owned by other shards are zero, so an
// element-wise SUM over shards reproduces the single-GPU buffer exactly.
#include "internal.h"

namespace xpcs {

constexpr int kSegWarps = 4;            // warp 0: ordered fp32 sums; warps 1..3: sum x, sum x^2
constexpr int kSegTile = 32;            // rows per tile
constexpr int kSegStages = 3;           // cp.async ring depth
constexpr int kSegPitch = kSegTile + 1; // padded: lane = delay reads down a row conflict free
constexpr int kSegTileFloats = 3 * 32 * kSegPitch;

struct SegArgs {
    const float *G2, *IP, *IF;
    const int *lseg_row_start;
    double *partials;
    int R_pad, T, nseg_local, seg_first, nseg_total;
};

__device__ __forceinline__ void cp_async_4(float *smem_dst, const float *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// One CTA per (static segment, chunk of 32 delays).  The [32 delays][32 rows] tiles of G2, IP
// and IF stream through a 3-deep cp.async ring (coalesced 128-byte row pieces in, padded pitch
// in shared memory).  Warp 0, lane = delay, folds every tile in row order: the three fp32 sums
// run in exactly the reference's order (corr.cpp:966-991).  Warps 1..3 accumulate, in fp64 and
// over the rows with (row - r0) % 3 == warp - 1, the sum and sum of squares of the per-pixel
// x = G2/(IP*IF) (NaN -> 0, corr.cpp:1066-1070); the three partial sums are combined in a fixed
// order, so the result depends on the segment's rows only (not on the GPU count).
__global__ void __launch_bounds__(kSegWarps * 32) k_segment_reduce(SegArgs a)
{
    extern __shared__ __align__(16) float seg_smem[];
    __shared__ double xs[kSegWarps - 1][2][32];
    const int seg = blockIdx.x;
    const int t0 = blockIdx.y * 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int r0 = a.lseg_row_start[seg], r1 = a.lseg_row_start[seg + 1];
    const int nt = min(32, a.T - t0);
    const int ntiles = (r1 - r0 + kSegTile - 1) / kSegTile;

    auto issue = [&](int tile) {
        if (tile < ntiles) {
            float *st = seg_smem + (size_t)(tile % kSegStages) * kSegTileFloats;
            const int base = r0 + tile * kSegTile;
            if (base + lane < r1) {
                for (int t = warp; t < nt; t += kSegWarps) {
                    const int64_t o = (int64_t)(t0 + t) * a.R_pad + base + lane;
                    cp_async_4(st + (0 * 32 + t) * kSegPitch + lane, a.G2 + o);
                    cp_async_4(st + (1 * 32 + t) * kSegPitch + lane, a.IP + o);
                    cp_async_4(st + (2 * 32 + t) * kSegPitch + lane, a.IF + o);
                }
            }
        }
        cp_async_commit();
    };

    for (int s = 0; s < kSegStages - 1; s++) issue(s);
    float sg = 0.0f, sp = 0.0f, sf = 0.0f;
    double sx = 0.0, sxx = 0.0;
    for (int tile = 0; tile < ntiles; tile++) {
        issue(tile + kSegStages - 1);
        cp_async_wait<kSegStages - 1>();
        __syncthreads();
        const float *st = seg_smem + (size_t)(tile % kSegStages) * kSegTileFloats;
        const int rows = min(kSegTile, r1 - r0 - tile * kSegTile);
        if (lane < nt) {
            const float *tg = st + (0 * 32 + lane) * kSegPitch;
            const float *tp = st + (1 * 32 + lane) * kSegPitch;
            const float *tf = st + (2 * 32 + lane) * kSegPitch;
            if (warp == 0) {
                for (int rr = 0; rr < rows; rr++) {
                    sg = __fadd_rn(sg, tg[rr]);
                    sp = __fadd_rn(sp, tp[rr]);
                    sf = __fadd_rn(sf, tf[rr]);
                }
            } else {
                int rr = (warp - 1) - (tile * kSegTile) % 3;  // (tile*32 + rr) % 3 == warp - 1
                if (rr < 0) rr += 3;
                for (; rr < rows; rr += 3) {
                    float x = __fdiv_rn(tg[rr], __fmul_rn(tp[rr], tf[rr]));
                    if (x != x) x = 0.0f;
                    sx += (double)x;
                    sxx += (double)x * (double)x;
                }
            }
        }
        __syncthreads();
    }
    if (warp > 0) {
        xs[warp - 1][0][lane] = sx;
        xs[warp - 1][1][lane] = sxx;
    }
    __syncthreads();
    if (warp == 0 && lane < nt) {
        const float cnt = (float)(r1 - r0);
        sg = __fdiv_rn(sg, cnt);
        sp = __fdiv_rn(sp, cnt);
        sf = __fdiv_rn(sf, cnt);
        const float g2 = __fdiv_rn(sg, __fmul_rn(sp, sf));
        const int64_t o = (int64_t)(a.seg_first + seg) * a.T + t0 + lane;
        const int64_t plane = (int64_t)a.nseg_total * a.T;
        a.partials[o] = (double)g2;
        a.partials[plane + o] = (xs[0][0][lane] + xs[1][0][lane]) + xs[2][0][lane];
        a.partials[2 * plane + o] = (xs[0][1][lane] + xs[1][1][lane]) + xs[2][1][lane];
    }
}

struct FinArgs {
    const double *partials;
    const int *seg_dq;        // [nseg_total] ascending
    const int *seg_pixels_n;  // [nseg_total]
    float *g2, *se;           // (T, Q)
    int T, Q, nseg_total;
};

__global__ void k_normalize_finish(FinArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y + 1;
    if (t >= a.T) return;
    const int64_t plane = (int64_t)a.nseg_total * a.T;
    float acc = 0.0f, cnt = 0.0f;
    double sx = 0.0, sxx = 0.0, n = 0.0;
    bool any = false;
    for (int s = 0; s < a.nseg_total; s++) {
        if (a.seg_dq[s] != q) continue;
        any = true;
        const float x = (float)a.partials[(int64_t)s * a.T + t];
        const bool nan = x != x;
        acc = __fadd_rn(acc, nan ? 0.0f : x);
        cnt = __fadd_rn(cnt, nan ? 0.0f : 1.0f);
        sx += a.partials[plane + (int64_t)s * a.T + t];
        sxx += a.partials[2 * plane + (int64_t)s * a.T + t];
        n += (double)a.seg_pixels_n[s];
    }
    if (!any) return;  // rows of absent dynamic bins stay zero (g2.setZero, corr.cpp:947)
    a.g2[(int64_t)t * a.Q + (q - 1)] = __fdiv_rn(acc, cnt);
    const double m2 = sxx - sx * sx / n;
    const float nf = (float)n;
    const float norm = __fdiv_rn((float)(m2 > 0.0 ? m2 : 0.0), nf);
    const float inv = __fdiv_rn(1.0f, nf);
    a.se[(int64_t)t * a.Q + (q - 1)] = __fmul_rn(sqrtf(inv), sqrtf(norm));
}

int launch_normalize_partials(xpcs_handle_s *h)
{
    const int64_t plane = (int64_t)h->nseg_total * h->T;
    h->partials_count = 3 * plane;
    int rc = ensure(h, h->d_partials, (size_t)(h->partials_count > 0 ? h->partials_count : 1), "partials");
    if (rc) return rc;
    cudaMemsetAsync(h->d_partials.p, 0, sizeof(double) * (size_t)h->partials_count, h->stream);
    const int nseg_local = h->seg_last - h->seg_first;
    if (nseg_local > 0 && h->T > 0) {
        SegArgs a{};
        a.G2 = h->d_G2.p;
        a.IP = h->d_IP.p;
        a.IF = h->d_IF.p;
        a.lseg_row_start = h->d_lseg_row_start.p;
        a.partials = h->d_partials.p;
        a.R_pad = h->R_pad;
        a.T = h->T;
        a.nseg_local = nseg_local;
        a.seg_first = h->seg_first;
        a.nseg_total = h->nseg_total;
        dim3 grid(nseg_local, (h->T + 31) / 32);
        const size_t smem = sizeof(float) * (size_t)kSegStages * kSegTileFloats;
        rc = check_cuda(h, cudaFuncSetAttribute(k_segment_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "segment reduce smem attr");
        if (rc) return rc;
        LaunchScope ls(h, "k_segment_reduce");
        k_segment_reduce<<<grid, kSegWarps * 32, smem, h->stream>>>(a);
    }
    return check_cuda(h, cudaGetLastError(), "k_segment_reduce");
}

int launch_normalize_finish(xpcs_handle_s *h, float *d_g2, float *d_se)
{
    cudaMemsetAsync(d_g2, 0, sizeof(float) * (size_t)h->T * h->Q, h->stream);
    cudaMemsetAsync(d_se, 0, sizeof(float) * (size_t)h->T * h->Q, h->stream);
    if (h->Q > 0 && h->T > 0) {
        FinArgs a{};
        a.partials = h->d_partials.p;
        a.seg_dq = h->d_seg_dq_all.p;
        a.seg_pixels_n = h->d_seg_npix_all.p;
        a.g2 = d_g2;
        a.se = d_se;
        a.T = h->T;
        a.Q = h->Q;
        a.nseg_total = h->nseg_total;
        dim3 grid((h->T + 63) / 64, h->Q);
        LaunchScope ls(h, "k_normalize_finish");
        k_normalize_finish<<<grid, 64, 0, h->stream>>>(a);
    }
    return check_cuda(h, cudaGetLastError(), "k_normalize_finish");
}

}  // namespace xpcs
