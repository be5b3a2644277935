// twotime.cu -- two-time correlation of one dynamic partition (reference corr.cpp:781-924,
// smoothing corr.cpp:433-560 / :1166-1305), the one dense contraction of the path and the one
// stage on the tensor cores.
//
// For the N pixels of the partition and F frames, with X[t][p] the filtered intensity:
//     sg[t]   = (1/N) sum_p X[t][p]                    (ComputeSGSymmetric, corr.cpp:1166-1226;
//                                                        "Average": its mean over t)
//     C[t1,t2] = (1/N) sum_p X[t1][p] X[t2][p] / (sg[t1] sg[t2])   for t2 >= t1, 0 below
//     g2full[d] = mean of the d-th diagonal, g2partials[d][w] windowed diagonal sums / wsize
// i.e. C = triu(X X^T) scaled in the epilogue: the reference divides every event by sg first
// (corr.cpp:516-536), which for symmetric smoothing factors out of the contraction, so the
// tensor cores see the raw photon counts -- exact in fp16, exact fp32 accumulation -- and the
// only rounding happens in the epilogue.
//
// k_twotime_gemm: one CTA per 128x256 tile of the upper triangle.  Operands are 128x64 (A) and 256x64 (B)
// fp16 tiles of the frame-major matrix Xt[F][Npad] (K-major for both), fetched by TMA
// (cp.async.bulk.tensor.2d, SWIZZLE_128B) through a 4-stage mbarrier ring; one thread issues
// tcgen05.mma (cta_group::1, kind::f16, M=128, N=256, K=16) with the fp32 accumulator in TMEM
// (256 columns); tcgen05.commit releases smem stages and finally signals the epilogue, whose
// four warps read TMEM with tcgen05.ld.32x32b, scale and store.  The wide tile moves 48 KB of operands
// per 4.2 MFLOP (87 flop/B against 64 for 128x128): the shared-memory fill out of L2, not the tensor
// pipe, is what bounds this contraction (K = the pixel count of the partition is long, 1024 k-blocks at
// 64k pixels).  The CTAs walk the tiles in super-blocks of 16 x 8 tiles (a host-built tile list), so that the
// ~148 CTAs in flight share 24 row panels of the operand in L2 instead of ~80 (round 1: 33 GB of DRAM
// reads for a 1.3 GB operand).
// Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer, 2..5 = epilogue.
// Float-valued rows (flat-field, averaging) are split x = hi + lo in fp16 and contracted in
// three passes (hi*hi + hi*lo + lo*hi) into the same accumulator (~2e-7 relative); so are integer
// counts above 2048, which fp16 no longer holds exactly.
// "StaticMap" smoothing (SmoothingStaticMap / ComputeSGStaticMap, corr.cpp:433-494, :1228-1305) divides
// every event by the sg of its STATIC partition, which does not factor out of the contraction: the
// operand is built from z = x / sg_s[t] (one fp32 division per event, as the reference does) and takes
// the three-pass path; the epilogue then only divides by N.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>

#include "internal.h"

namespace xpcs {

constexpr int kTtBM = 128, kTtBN = 256, kTtBK = 64, kTtStages = 4;
constexpr int kTtTileA = kTtBM * kTtBK * 2;      // 16 KiB
constexpr int kTtTileB = kTtBN * kTtBK * 2;      // 32 KiB
constexpr int kTtStageBytes = kTtTileA + kTtTileB;
constexpr int kTtThreads = 192;
constexpr int kTtTmemCols = 256;
constexpr size_t kTtSmemBytes = (size_t)kTtStages * kTtStageBytes + 1024 /*align*/ + 256 /*barriers*/;

// ---- PTX wrappers --------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row atoms of 1024 B (SBO = 1024),
// LBO = 16 B (unused for swizzled K-major), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16: A = B = fp16 (0), D = fp32 (1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kTtIdesc = (1u << 4) | ((uint32_t)(kTtBN >> 3) << 17) | ((uint32_t)(kTtBM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kTtIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TtGemmArgs {
    const int2 *tiles;  // (row tile of 128, column tile of 256) of every CTA, in L2-friendly order
    float *C;           // [F][F] row-major, pre-zeroed
    const float *sg;    // [F] (or [1] when sg_scalar)
    int F, kblocks, npass, sg_scalar, use_sg;
    float npix;
    float unscale;      // 1 / op_scale^2 (a power of two): the operand was scaled to stay inside fp16
};

__global__ void __launch_bounds__(kTtThreads, 1)
k_twotime_gemm(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
               const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo, TtGemmArgs a)
{
    // only tiles that touch the upper triangle are listed (the lower one stays zero, corr.cpp:826: k starts at j)
    const int2 tile = a.tiles[blockIdx.x];
    const int mt = tile.x, nt = tile.y;
    extern __shared__ unsigned char tt_smem_raw[];
    const uint32_t base = (smem_u32(tt_smem_raw) + 1023u) & ~1023u;
    const uint32_t tiles = base;                                    // [stage]: A 16 KiB, B 32 KiB
    const uint32_t bars = base + kTtStages * kTtStageBytes;         // full[4], empty[4], tmem_full, tmem_slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kTtStages + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * kTtStages);
    const uint32_t tmem_slot = bars + 8u * (2 * kTtStages + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_kb = a.kblocks * a.npass;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kTtStages; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTtTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            for (int kb = 0; kb < total_kb; kb++) {
                const int s = kb % kTtStages;
                const uint32_t ph = (uint32_t)(kb / kTtStages) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                const int pass = kb / a.kblocks, kk = kb - pass * a.kblocks;
                const CUtensorMap *ma = pass == 2 ? &mapA_lo : &mapA_hi;
                const CUtensorMap *mb = pass == 1 ? &mapB_lo : &mapB_hi;
                mbar_expect_tx(full_bar(s), kTtStageBytes);
                tma_load_2d(tiles + (uint32_t)s * kTtStageBytes, ma, kk * kTtBK, mt * kTtBM, full_bar(s));
                tma_load_2d(tiles + (uint32_t)s * kTtStageBytes + kTtTileA, mb, kk * kTtBK, nt * kTtBN, full_bar(s));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            for (int kb = 0; kb < total_kb; kb++) {
                const int s = kb % kTtStages;
                const uint32_t ph = (uint32_t)(kb / kTtStages) & 1u;
                mbar_wait(full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = tiles + (uint32_t)s * kTtStageBytes;
                const uint32_t sb = sa + kTtTileA;
#pragma unroll
                for (int k = 0; k < kTtBK / 16; k++) {
                    // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle atom
                    umma_f16(tmem_base, umma_desc_sw128(sa + 32u * k), umma_desc_sw128(sb + 32u * k),
                             (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(empty_bar(s));  // frees the stage once these MMAs have read it
            }
            umma_commit(tmem_full_bar);
        }
    } else {  // ===== epilogue: warps 2..5 own TMEM lane quadrants 2,3,0,1 =====
        const int q = warp & 3;
        mbar_wait(tmem_full_bar, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = mt * kTtBM + q * 32 + lane;
        float s_row = 1.0f;
        if (a.use_sg && row < a.F) s_row = a.sg_scalar ? a.sg[0] : a.sg[row];
        for (int c = 0; c < kTtBN / 32; c++) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            const int col0 = nt * kTtBN + c * 32;
            if (row < a.F) {
                float *dst = a.C + (size_t)row * a.F + col0;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int col = col0 + j;
                    if (col < a.F && col >= row) {
                        float x = __uint_as_float(v[j]);
                        if (x != 0.0f) {
                            if (a.use_sg) {
                                const float s_col = a.sg_scalar ? s_row : a.sg[col];
                                x = __fdiv_rn(x, __fmul_rn(s_row, s_col));
                            }
                            x = __fdiv_rn(__fmul_rn(x, a.unscale), a.npix);
                        }
                        dst[j] = x;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTtTmemCols));
    }
}

// ---- operand construction from the pixel-major store -----------------------------------
struct TtBuildArgs {
    const void *store;
    const int64_t *slice_base;
    const int *row_len;
    __half *xt_hi, *xt_lo;        // [F][npad]; xt_lo nullptr = single pass (exact small integers)
    unsigned int *sg_int;         // [sg_rows][F] integer column sums (packed rows)
    float *sg_f;                  // [sg_rows][F] float column sums
    int row0, row1, npad, F, n_slices;
    // static-map smoothing: rows belong to sg row (segment index - seg0); symmetric: everything is sg row 0
    const int *lseg_row_start;    // [local segments + 1]
    int seg0, nseg, per_segment;
    int do_sum, do_write;         // accumulate the column sums / write the operand
    const float *sg_div;          // static map, operand pass: divide by sg_div[row * (scalar ? 1 : F) + t]
    int sg_scalar;
    unsigned int *max_bits;       // sum pass: bit pattern of the largest value seen (values are >= 0)
    float op_scale;               // power of two
};

template <int KIND>
__global__ void k_twotime_build(TtBuildArgs a)
{
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= a.n_slices) return;
    const int r = s * kSlice + lane;
    if (r < a.row0 || r >= a.row1) return;
    const int n = a.row_len[r];
    const int p = r - a.row0;
    int sgrow = 0;
    if (a.per_segment) {  // last segment of [seg0, seg0 + nseg) that starts at or before r
        int lo = a.seg0, hi = a.seg0 + a.nseg;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.lseg_row_start[mid] <= r) lo = mid;
            else hi = mid;
        }
        sgrow = lo - a.seg0;
    }
    const size_t sgbase = (size_t)sgrow * a.F;
    unsigned int mb = 0u;
    for (int j = 0; j < n; j++) {
        int t;
        float x;
        unsigned c = 0;
        if (KIND == kPacked) {
            const uint32_t w = (reinterpret_cast<const uint32_t *>(a.store) + a.slice_base[s] + lane)[(int64_t)j * kSlice];
            t = (int)(w >> kCountBits);
            c = w & ((1u << kCountBits) - 1u);
            x = (float)c;
        } else {
            const unsigned long long w = (reinterpret_cast<const unsigned long long *>(a.store) + a.slice_base[s] + lane)[(int64_t)j * kSlice];
            t = (int)(w >> 32);
            x = __uint_as_float((uint32_t)w);
        }
        if (t >= a.F) continue;
        if (a.do_sum) {
            if (KIND == kPacked) atomicAdd(a.sg_int + sgbase + t, c);
            else atomicAdd(a.sg_f + sgbase + t, x);
            mb = max(mb, __float_as_uint(fmaxf(x, 0.0f)));
        }
        if (a.do_write) {
            if (a.sg_div) x = __fdiv_rn(x, a.sg_div[a.sg_scalar ? (size_t)sgrow : sgbase + t]);  // corr.cpp:470-478
            x = __fmul_rn(x, a.op_scale);
            const __half hi = __float2half_rn(x);
            a.xt_hi[(size_t)t * a.npad + p] = hi;
            if (a.xt_lo) a.xt_lo[(size_t)t * a.npad + p] = __float2half_rn(__fsub_rn(x, __half2float(hi)));
        }
    }
    if (a.do_sum && mb) atomicMax(a.max_bits, mb);
}

// sg[t] = column sum / N (corr.cpp:1207-1215); "Average": one value, the fp32 mean over t in
// frame order (corr.cpp:1217-1224)
// blockIdx.y = sg row (one for symmetric smoothing, one per static partition for StaticMap)
__global__ void k_twotime_sg(const unsigned int *sg_int, const float *sg_f, float *sg, const float *npix_of_row, int F,
                             int packed)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t o = (size_t)blockIdx.y * F + t;
    if (t < F) sg[o] = __fdiv_rn(packed ? (float)sg_int[o] : sg_f[o], npix_of_row[blockIdx.y]);
}
__global__ void k_twotime_sg_average(const float *sg, float *sg_avg, int F)
{
    if (threadIdx.x == 0) {
        const float *row = sg + (size_t)blockIdx.x * F;
        float acc = 0.0f;
        for (int t = 0; t < F; t++) acc = __fadd_rn(acc, row[t]);
        sg_avg[blockIdx.x] = __fdiv_rn(acc, (float)F);
    }
}

// Diagonal statistics of C (corr.cpp:842-866): thread = diagonal d, rows in chunks; adjacent
// threads read adjacent addresses.  Sums in fp64, combined with atomics.
__global__ void k_twotime_diag(const float *__restrict__ C, double *__restrict__ full, double *__restrict__ part,
                               int F, int wsize, int partials, int rows_per_block)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= F) return;
    const int x0 = blockIdx.y * rows_per_block;
    const int x1 = min(F - d, x0 + rows_per_block);
    double acc = 0.0, wacc = 0.0;
    int win = -1;
    for (int x = x0; x < x1; x++) {
        const float v = C[(size_t)x * F + x + d];
        acc += (double)v;
        if (d < wsize) {
            const int w = x / wsize;
            if (w != win) {
                if (win >= 0 && win < partials) atomicAdd(part + (size_t)d * partials + win, wacc);
                win = w;
                wacc = 0.0;
            }
            wacc += (double)v;
        }
    }
    if (d < wsize && win >= 0 && win < partials) atomicAdd(part + (size_t)d * partials + win, wacc);
    if (x1 > x0) atomicAdd(full + d, acc);
}
__global__ void k_twotime_diag_finish(const double *full, const double *part, float *g2full, float *g2part, int F,
                                      int wsize, int partials)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F) g2full[i] = __fdiv_rn((float)full[i], (float)(F - i));
    if (i < wsize * partials) g2part[i] = __fdiv_rn((float)part[i], (float)wsize);
}

// ---- host side ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_operand_map(xpcs_handle_s *h, CUtensorMap *map, const __half *ptr, int F, int npad, int box_rows)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || !p) return fail(h, XPCS_E_CUDA, "cuTensorMapEncodeTiled not available (%s)", cudaGetErrorString(e));
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)npad, (cuuint64_t)F};
    const cuuint64_t strides[1] = {(cuuint64_t)npad * sizeof(__half)};
    const cuuint32_t box[2] = {(cuuint32_t)kTtBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)ptr, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, XPCS_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return XPCS_OK;
}

int launch_twotime(xpcs_handle_s *h, int qbin, int wsize, int method, int average, float *C, float *g2full,
                   float *g2partials, float *sg, int *sg_rows_out)
{
    const int F = h->prm.frames;
    if (method < 0 || method > 2) return fail(h, XPCS_E_ARG, "two-time: smoothing method %d unknown (0 = none, 1 = symmetric, 2 = static map)", method);
    if (wsize <= 0) return fail(h, XPCS_E_ARG, "two-time: twotime2onetime_window_size must be > 0 (corr.cpp:796 divides by it)");
    if (h->prm.shard_count != 1) return fail(h, XPCS_E_ARG, "two-time runs one dynamic partition per GPU: create the handle with shard_count = 1");
    // rows of the dynamic partition: contiguous, because rows are sorted by (dq, sq, pixel)
    int row0 = -1, row1 = -1, seg0 = -1, nseg = 0;
    for (int s = h->seg_first; s < h->seg_last; s++)
        if (h->seg_dq[s] == qbin) {
            if (row0 < 0) {
                row0 = h->lseg_row_start[s - h->seg_first];
                seg0 = s - h->seg_first;
            }
            row1 = h->lseg_row_start[s - h->seg_first + 1];
            nseg++;
        }
    if (row0 < 0) return fail(h, XPCS_E_ARG, "two-time: dynamic partition %d has no pixels", qbin);
    const int N = row1 - row0;
    const int npad = (N + kTtBK - 1) / kTtBK * kTtBK;
    const int partials = (F - wsize) / wsize > 0 ? (F - wsize) / wsize : 0;
    const bool packed = h->kind == kPacked;
    const bool static_map = method == 2;
    const int sg_rows = static_map ? nseg : 1;
    if (sg_rows_out) *sg_rows_out = sg_rows;
    // the fp16 operand is exact for integer counts up to 2048 only; anything else takes hi + lo and three passes
    const bool split = !packed || static_map || h->max_count > 2048;
    float op_scale = 1.0f;  // static map: set once sg is known (below)
    int rc;
    const size_t xt_elems = (size_t)F * npad;
    const size_t sg_elems = (size_t)sg_rows * F;
    if ((rc = ensure(h, h->d_tt_hi, xt_elems, "two-time operand"))) return rc;
    if (split && (rc = ensure(h, h->d_tt_lo, xt_elems, "two-time operand (low part)"))) return rc;
    if ((rc = ensure(h, h->d_tt_C, (size_t)F * F, "two-time matrix"))) return rc;
    if ((rc = ensure(h, h->d_tt_sg, 2 * sg_elems + 2 * (size_t)sg_rows + 8, "two-time sg"))) return rc;
    if ((rc = ensure(h, h->d_tt_sgint, sg_elems + 1, "two-time sg sums"))) return rc;
    if ((rc = ensure(h, h->d_tt_diag, (size_t)F + (size_t)wsize * (partials > 0 ? partials : 1), "two-time diagonals"))) return rc;
    if ((rc = ensure(h, h->d_tt_out, (size_t)F + (size_t)wsize * (partials > 0 ? partials : 1), "two-time diagonal means"))) return rc;
    cudaStream_t st = h->stream;
    cudaMemsetAsync(h->d_tt_hi.p, 0, xt_elems * sizeof(__half), st);
    if (split) cudaMemsetAsync(h->d_tt_lo.p, 0, xt_elems * sizeof(__half), st);
    cudaMemsetAsync(h->d_tt_C.p, 0, (size_t)F * F * sizeof(float), st);
    cudaMemsetAsync(h->d_tt_sg.p, 0, (2 * sg_elems + 2 * (size_t)sg_rows + 8) * sizeof(float), st);
    cudaMemsetAsync(h->d_tt_sgint.p, 0, (sg_elems + 1) * sizeof(unsigned int), st);
    cudaMemsetAsync(h->d_tt_diag.p, 0, ((size_t)F + (size_t)wsize * (partials > 0 ? partials : 1)) * sizeof(double), st);

    float *d_sg = h->d_tt_sg.p;                          // [sg_rows][F] per-frame sg
    float *d_sg_f = h->d_tt_sg.p + sg_elems;             // [sg_rows][F] float column sums
    float *d_sg_avg = h->d_tt_sg.p + 2 * sg_elems;       // [sg_rows]
    float *d_npix = d_sg_avg + sg_rows;                  // [sg_rows] pixels behind every sg row
    std::vector<float> npix((size_t)sg_rows, (float)N);  // (stays alive until the synchronisation below)
    if (static_map)
        for (int k = 0; k < nseg; k++) npix[(size_t)k] = (float)(h->lseg_row_start[seg0 + k + 1] - h->lseg_row_start[seg0 + k]);
    if ((rc = check_cuda(h, cudaMemcpyAsync(d_npix, npix.data(), sizeof(float) * (size_t)sg_rows, cudaMemcpyHostToDevice, st), "sg pixel counts")))
        return rc;
    TtBuildArgs b{};
    b.store = h->d_store.p;
    b.slice_base = h->d_slice_base.p;
    b.row_len = h->d_row_len.p;
    b.xt_hi = (__half *)h->d_tt_hi.p;
    b.xt_lo = split ? (__half *)h->d_tt_lo.p : nullptr;
    b.sg_int = h->d_tt_sgint.p;
    b.sg_f = d_sg_f;
    b.row0 = row0;
    b.row1 = row1;
    b.npad = npad;
    b.F = F;
    b.n_slices = h->n_slices;
    b.lseg_row_start = h->d_lseg_row_start.p;
    b.seg0 = seg0;
    b.nseg = nseg;
    b.per_segment = static_map ? 1 : 0;
    b.op_scale = 1.0f;
    b.max_bits = h->d_tt_sgint.p + sg_elems;
    const int wpb = 8;
    auto run_build = [&](int do_sum, int do_write, const float *div) {
        b.do_sum = do_sum;
        b.do_write = do_write;
        b.sg_div = div;
        b.sg_scalar = average ? 1 : 0;
        LaunchScope ls(h, "k_twotime_build");
        if (packed) k_twotime_build<kPacked><<<(h->n_slices + wpb - 1) / wpb, wpb * 32, 0, st>>>(b);
        else k_twotime_build<kFloat><<<(h->n_slices + wpb - 1) / wpb, wpb * 32, 0, st>>>(b);
    };
    run_build(1, static_map ? 0 : 1, nullptr);  // column sums (and, when sg factors out, the operand itself)
    {
        LaunchScope ls(h, "k_twotime_sg");
        k_twotime_sg<<<dim3((F + 255) / 256, sg_rows), 256, 0, st>>>(h->d_tt_sgint.p, d_sg_f, d_sg, d_npix, F, packed ? 1 : 0);
    }
    if (average) {
        LaunchScope ls(h, "k_twotime_sg_average");
        k_twotime_sg_average<<<sg_rows, 32, 0, st>>>(d_sg, d_sg_avg, F);
    }
    if (static_map) {
        // z = x / sg_s can leave the fp16 range: per frame z <= N_s (x over the mean of a sum that contains it),
        // with the "Average" filter z <= max x / min sg.  The operand is scaled by a power of two, undone in the epilogue.
        unsigned int mb = 0;
        std::vector<float> avg((size_t)sg_rows, 1.0f);
        cudaMemcpyAsync(&mb, h->d_tt_sgint.p + sg_elems, sizeof(mb), cudaMemcpyDeviceToHost, st);
        if (average) cudaMemcpyAsync(avg.data(), d_sg_avg, sizeof(float) * (size_t)sg_rows, cudaMemcpyDeviceToHost, st);
        if ((rc = check_cuda(h, cudaStreamSynchronize(st), "two-time sg"))) return rc;
        float maxval;
        memcpy(&maxval, &mb, sizeof(float));
        double bound = 1.0;
        for (int k = 0; k < sg_rows; k++) {
            if (average) {
                if (avg[(size_t)k] > 0.0f) bound = std::max(bound, (double)maxval / (double)avg[(size_t)k]);
            } else bound = std::max(bound, (double)npix[(size_t)k]);
        }
        while (bound * op_scale > 16384.0 && op_scale > 1e-12f) op_scale *= 0.5f;
        b.op_scale = op_scale;
        run_build(0, 1, average ? d_sg_avg : d_sg);  // operand z = x / sg of the pixel's static partition
    }
    CUtensorMap mapA_hi, mapA_lo, mapB_hi, mapB_lo;
    const __half *p_hi = (const __half *)h->d_tt_hi.p, *p_lo = (const __half *)(split ? h->d_tt_lo.p : h->d_tt_hi.p);
    if ((rc = make_operand_map(h, &mapA_hi, p_hi, F, npad, kTtBM))) return rc;
    if ((rc = make_operand_map(h, &mapA_lo, p_lo, F, npad, kTtBM))) return rc;
    if ((rc = make_operand_map(h, &mapB_hi, p_hi, F, npad, kTtBN))) return rc;
    if ((rc = make_operand_map(h, &mapB_lo, p_lo, F, npad, kTtBN))) return rc;
    // tile list: super-blocks of 16 x 8 tiles (2048 x 2048 frames) row by row, tiles row-major inside; a tile is
    // listed when its last column reaches its first row
    std::vector<int2> tl;
    {
        const int tm = (F + kTtBM - 1) / kTtBM, tnn = (F + kTtBN - 1) / kTtBN;
        const int SM_ = 16, SN_ = 8;
        for (int bm = 0; bm < tm; bm += SM_)
            for (int bn = 0; bn < tnn; bn += SN_)
                for (int mt = bm; mt < std::min(tm, bm + SM_); mt++)
                    for (int nt = bn; nt < std::min(tnn, bn + SN_); nt++)
                        if (nt * kTtBN + kTtBN - 1 >= mt * kTtBM) tl.push_back(make_int2(mt, nt));
    }
    if ((rc = ensure(h, h->d_tt_tiles, tl.size() * 2 + 2, "two-time tile list"))) return rc;
    if ((rc = check_cuda(h, cudaMemcpyAsync(h->d_tt_tiles.p, tl.data(), sizeof(int2) * tl.size(), cudaMemcpyHostToDevice, st), "tile list")))
        return rc;
    {
        TtGemmArgs g{};
        g.tiles = reinterpret_cast<const int2 *>(h->d_tt_tiles.p);
        g.C = h->d_tt_C.p;
        g.sg = average ? d_sg_avg : d_sg;
        g.F = F;
        g.kblocks = npad / kTtBK;
        g.npass = split ? 3 : 1;
        g.sg_scalar = average ? 1 : 0;
        g.use_sg = method == 1 ? 1 : 0;
        g.npix = (float)N;
        g.unscale = 1.0f / (op_scale * op_scale);
        rc = check_cuda(h, cudaFuncSetAttribute(k_twotime_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTtSmemBytes),
                        "two-time smem attr");
        if (rc) return rc;
        LaunchScope ls(h, "k_twotime_gemm");
        k_twotime_gemm<<<(unsigned)tl.size(), kTtThreads, kTtSmemBytes, st>>>(mapA_hi, mapA_lo, mapB_hi, mapB_lo, g);
    }
    if ((rc = check_cuda(h, cudaGetLastError(), "k_twotime_gemm"))) return rc;
    double *d_full = h->d_tt_diag.p;
    double *d_part = h->d_tt_diag.p + F;
    float *d_g2full = h->d_tt_out.p;
    float *d_g2part = h->d_tt_out.p + F;
    {
        const int rows_per_block = 256;
        LaunchScope ls(h, "k_twotime_diag");
        k_twotime_diag<<<dim3((F + 127) / 128, (F + rows_per_block - 1) / rows_per_block), 128, 0, st>>>(
            h->d_tt_C.p, d_full, d_part, F, wsize, partials, rows_per_block);
    }
    {
        const int n = F > wsize * partials ? F : wsize * partials;
        LaunchScope ls(h, "k_twotime_diag_finish");
        k_twotime_diag_finish<<<(n + 255) / 256, 256, 0, st>>>(d_full, d_part, d_g2full, d_g2part, F, wsize, partials);
    }
    if (C) cudaMemcpyAsync(C, h->d_tt_C.p, (size_t)F * F * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (g2full) cudaMemcpyAsync(g2full, d_g2full, (size_t)F * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (g2partials && partials > 0)
        cudaMemcpyAsync(g2partials, d_g2part, (size_t)wsize * partials * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (sg) {
        const size_t n = (size_t)sg_rows * (average ? 1 : (size_t)F);
        if (method != 0) cudaMemcpyAsync(sg, average ? d_sg_avg : d_sg, n * sizeof(float), cudaMemcpyDeviceToHost, st);
        else for (size_t i = 0; i < n; i++) sg[i] = 1.0f;
    }
    return check_cuda(h, cudaStreamSynchronize(st), "two-time");
}

}  // namespace xpcs
