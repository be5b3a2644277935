// normalize.cu -- q-bin normalisation: static-partition means, normalised g2 per dynamic
// bin and its standard error.  Replaces Corr::normalizeG2s (reference corr.cpp:927-1091).
//
// Rows are stored in (dq, sq, pixel) order, so every static segment is a contiguous row
// range of the [T][R_pad] correlator arrays.
//  k_segment_reduce  one warp per (segment, chunk of 32 delays).  The rows of the segment are
//                    streamed in tiles of 64 with coalesced loads (lanes along rows) into a
//                    padded shared tile; lane = delay then folds the tile in row order, so the
//                    three fp32 sums run in exactly the reference's order
//                    (corr.cpp:966-991: sequential += over the pixels of a static bin) at full
//                    memory efficiency.  It also accumulates, in fp64, the sum and sum of
//                    squares of the per-pixel x = G2/(IP*IF) (NaN -> 0, corr.cpp:1066-1070).
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

constexpr int kTileRows = 64;
constexpr int kTilePitch = kTileRows + 1;

struct SegArgs {
    const float *G2, *IP, *IF;
    const int *lseg_row_start;
    double *partials;
    int R_pad, T, nseg_local, seg_first, nseg_total;
};

__global__ void __launch_bounds__(32) k_segment_reduce(SegArgs a)
{
    __shared__ float tg[32 * kTilePitch], tp[32 * kTilePitch], tf[32 * kTilePitch];
    const int seg = blockIdx.x;
    const int t0 = blockIdx.y * 32;
    const int lane = threadIdx.x;
    const int r0 = a.lseg_row_start[seg], r1 = a.lseg_row_start[seg + 1];
    const int nt = min(32, a.T - t0);
    float sg = 0.0f, sp = 0.0f, sf = 0.0f;
    double sx = 0.0, sxx = 0.0;
    for (int base = r0; base < r1; base += kTileRows) {
        const int rows = min(kTileRows, r1 - base);
        __syncwarp();
        for (int t = 0; t < nt; t++) {
            const int64_t o = (int64_t)(t0 + t) * a.R_pad + base;
#pragma unroll
            for (int k = 0; k < kTileRows / 32; k++) {
                const int rr = k * 32 + lane;
                if (rr < rows) {
                    tg[t * kTilePitch + rr] = a.G2[o + rr];
                    tp[t * kTilePitch + rr] = a.IP[o + rr];
                    tf[t * kTilePitch + rr] = a.IF[o + rr];
                }
            }
        }
        __syncwarp();
        if (lane < nt) {
            for (int rr = 0; rr < rows; rr++) {
                const float g = tg[lane * kTilePitch + rr];
                const float p = tp[lane * kTilePitch + rr];
                const float f = tf[lane * kTilePitch + rr];
                sg = __fadd_rn(sg, g);
                sp = __fadd_rn(sp, p);
                sf = __fadd_rn(sf, f);
                float x = __fdiv_rn(g, __fmul_rn(p, f));
                if (x != x) x = 0.0f;
                sx += (double)x;
                sxx += (double)x * (double)x;
            }
        }
    }
    if (lane < nt) {
        const float cnt = (float)(r1 - r0);
        sg = __fdiv_rn(sg, cnt);
        sp = __fdiv_rn(sp, cnt);
        sf = __fdiv_rn(sf, cnt);
        const float g2 = __fdiv_rn(sg, __fmul_rn(sp, sf));
        const int64_t o = (int64_t)(a.seg_first + seg) * a.T + t0 + lane;
        const int64_t plane = (int64_t)a.nseg_total * a.T;
        a.partials[o] = (double)g2;
        a.partials[plane + o] = sx;
        a.partials[2 * plane + o] = sxx;
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
        LaunchScope ls(h, "k_segment_reduce");
        k_segment_reduce<<<grid, 32, 0, h->stream>>>(a);
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
