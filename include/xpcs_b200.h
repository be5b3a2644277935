/*
 * xpcs_b200.h -- C-ABI of the B200-native XPCS correlation hot path.
 *
 * The reference (AdvancedPhotonSource/xpcs-eigen) has no FFI; its seams for this path
 * are two abstract C++ classes and one static class, all reading a Configuration
 * singleton.  Each entry point below names the reference
 * interface it replaces (file:line relative to the reference root).  The host program
 * xpcs-eigen_b200/host (our `corr`) and the ctypes layer xpcs-eigen_b200/cabi.py bind
 * exactly these symbols; INTEGRATION.md shows the binding a reference maintainer adds.
 *
 * Conventions: plain C, opaque handle, int status (0 = ok, negative = XPCS_E_*),
 * xpcs_last_error() for the message, no exceptions cross the boundary, one handle per
 * GPU, thread-compatible (not thread-safe).  Host buffers are caller-owned; device
 * memory is owned by the handle unless an entry point says "device pointer".
 * Layouts are the reference's: tau-major [tau][pixel] for G2/IP/IF (main.cpp:193-201,
 * corr.cpp:395), (T, Q) for norm-0-g2 / norm-0-stderr (h5_result.cpp:80-81),
 * [2][F] for frameSum with row 0 = 1..F (sparse_filter.cpp:189-190).
 * There is NO CPU fallback: every compute entry point fails with XPCS_E_CUDA when no
 * sm_100 device is usable.
 */
#ifndef XPCS_B200_H
#define XPCS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XPCS_OK 0
#define XPCS_E_ARG (-1)     /* bad argument / inconsistent configuration          */
#define XPCS_E_CUDA (-2)    /* CUDA runtime or launch failure, or no usable device */
#define XPCS_E_STATE (-3)   /* call sequence violated (e.g. multitau before ingest) */
#define XPCS_E_NOMEM (-4)

/* compat_flags */
#define XPCS_COMPAT_STALE_TAIL 1u /* reproduce the reference's dropped G2 pairs: the binary
                                     search of corr.cpp:406 runs over the un-shrunk index
                                     vector.  Off = exact sums.            */
#define XPCS_COMPAT_LATE_WINDOW 2u /* static windows as the Rigaku reader counts them
                                     (io/rigaku.cpp:190-193: the window number moves on AFTER the
                                     frames 1w, 2w, ...), not as the Filter stage does
                                     (sparse_filter.cpp:160-162: before them): frame t > 0
                                     belongs to window (t-1)/w.                             */
#define XPCS_FLAG_LANE_MULTITAU 0x100u /* diagnostics: run the lane-per-row multi-tau kernel on
                                          every slice instead of the warp-per-row one (both are
                                          exact on integer counts; used by the parity tests to
                                          cross-check the two).                              */
#define XPCS_FLAG_SCALAR_DENSE 0x200u  /* diagnostics: run the scalar dense-filter kernel (the one
                                          used for detectors whose pixel count is not a multiple
                                          of 8) instead of the vectorised one.                */

typedef struct xpcs_handle_s *xpcs_handle;

/* What Configuration::init hands to the filters and to Corr (configuration.cpp:80-242);
 * here an explicit POD instead of a process-wide singleton. */
typedef struct XpcsParams {
    int32_t struct_size;           /* = sizeof(XpcsParams), ABI guard                         */
    int32_t width, height;         /* detector x_dimension / y_dimension (:94-95); P = w*h     */
    int32_t frames;                /* getFrameTodoCount() (:594-601): output frames F          */
    int32_t delays_per_level;      /* <entry>/delays_per_level (:146)                          */
    int32_t stride_frames;         /* <entry>/stride_frames (:154)                             */
    int32_t avg_frames;            /* <entry>/avg_frames (:155)                                */
    int32_t static_window;         /* <entry>/static_mean_window_size (:201)                   */
    int32_t normalize_by_framesum; /* <entry>/normalize_by_framesum (:158-162)                 */
    uint32_t compat_flags;         /* XPCS_COMPAT_*                                            */
    float lld, sigma;              /* <entry>/lld, sigma (:176-177); used with dark frames     */
    const int32_t *dqmap;          /* [P] row-major dynamic partition ids, 0 = masked (:97)    */
    const int32_t *sqmap;          /* [P] static partition ids (:98)                           */
    const double *flatfield;       /* [P] or NULL = all ones (:203-214)                        */
    int32_t shard_index;           /* this handle's pixel shard, 0 <= shard_index < shard_count */
    int32_t shard_count;           /* 1 = whole detector.  Shards are contiguous ranges of the  */
                                   /* (dq, sq, pixel)-sorted pixel list cut at static-bin       */
                                   /* boundaries and balanced by pixel count.    */
    int64_t reserve_events;        /* optional capacity hint for the device event store         */
} XpcsParams;

typedef struct XpcsInfo {
    int32_t n_delays;        /* T                                                            */
    int32_t max_level;
    int32_t n_static;        /* S = getTotalStaticPartitions()                               */
    int32_t n_dynamic;       /* Q = getTotalDynamicPartitions()                              */
    int32_t n_segments;      /* surviving (dq, sq) map entries, all shards                   */
    int32_t n_rows;          /* unmasked pixels owned by this shard                          */
    int32_t n_rows_total;    /* unmasked pixels of the whole detector                        */
    int32_t raw_frames_seen; /* raw frames pushed so far                                     */
    int64_t events_pushed;   /* raw events received                                          */
    int64_t events_stored;   /* events in the pixel-major store after finish_ingest          */
    int64_t store_words;     /* padded words of the pixel-major store                        */
    int32_t value_kind;      /* 0 = packed integer counts (exact), 1 = float values          */
    int32_t max_row_events;  /* longest pixel row                                            */
} XpcsInfo;

/* ---- schedule: Corr::calculateLevelMax / Corr::delaysPerLevel (corr.cpp:1133-1160) ---- */
int xpcs_level_max(int frames, int delays_per_level);
/* writes up to cap (level, tau) pairs, returns T */
int xpcs_delay_schedule(int frames, int delays_per_level, int32_t *level, int32_t *tau, int cap);

/* ---- pixel sharding (host only: callable on a box without a GPU) ---- */
typedef struct XpcsShardPlan {
    int32_t n_static, n_dynamic; /* S, Q                                                          */
    int32_t n_segments;          /* surviving (dq, sq) map entries of the whole detector          */
    int32_t seg_first, seg_last; /* this shard owns global segments [seg_first, seg_last)          */
    int32_t n_rows;              /* unmasked pixels owned by this shard                            */
    int32_t n_rows_total;
    int32_t n_delays;            /* T                                                              */
} XpcsShardPlan;
/* The partition maps of Configuration::BuildQMap (configuration.cpp:244-381) and the shard
 * (params->shard_index of shard_count) xpcs_create would build, without touching a device.
 * row_pixels (nullable, capacity cap): detector pixel of every owned row in store order. */
int xpcs_plan_shard(const XpcsParams *params, XpcsShardPlan *plan, int32_t *row_pixels, int64_t cap);

/* ---- lifetime ---- */
/* replaces: Configuration::BuildQMap (configuration.cpp:244-381) + the constructors of
 * SparseFilter / DenseFilter (sparse_filter.cpp:63-109, dense_filter.cpp:64-115).       */
int xpcs_create(const XpcsParams *params, int device, xpcs_handle *out);
void xpcs_destroy(xpcs_handle h);
const char *xpcs_last_error(xpcs_handle h); /* h may be NULL: error of the last failed create */
int xpcs_get_info(xpcs_handle h, XpcsInfo *info);
/* detector pixel index of every row this shard owns, in store order: (dq, sq, pixel) sorted
 * (the iteration order of Configuration::getBinMaps(), configuration.cpp:262-296); out[n_rows] */
int xpcs_get_row_pixels(xpcs_handle h, int32_t *out);
/* run every kernel of this handle on the caller's CUDA stream (cudaStream_t); NULL = own stream */
int xpcs_set_stream(xpcs_handle h, void *cuda_stream);
/* forget all ingested frames and results, keep maps, dark image and allocations */
int xpcs_reset(xpcs_handle h);

/* ---- Filter stage ---- */
/* replaces: DarkImage::DarkImage/Compute (data_structure/dark_image.cpp:59-106) fed by
 * reader->NextFrames(darks) (main.cpp:227-239).  frames = [n][P] raw int16, host memory. */
int xpcs_set_dark(xpcs_handle h, const int16_t *frames, int n);
/* DarkImage::dark_avg()/dark_std() for --darkout (main.cpp:459-477); [P] each, nullable    */
int xpcs_get_dark(xpcs_handle h, double *avg, double *std);

/* replaces: Imm::NextFrames for compressed files (io/imm.cpp:70-118) + SparseFilter::Apply
 * (filter/sparse_filter.cpp:115-193) for `nframes` consecutive RAW frames.
 * idx/val = the concatenated payloads (int32 pixel index, int16 value) of those frames,
 * frame_offsets[nframes+1] = event offsets of each frame relative to idx/val,
 * clock/ticks = Header::elapsed / Header::corecotick per raw frame (nullable).
 * Host pointers; idx/val must stay valid and unmodified until xpcs_finish_ingest returns
 * unless the push was pipelined (then it has been consumed on return).  With pinned memory
 * the copies are asynchronous, and a first push of >= 4 Mi events of plain photon counts
 * (no flat-field, stride, averaging or frame-sum normalisation) is PIPELINED: cut into up to
 * 4 chunks of frames, chunk k ingested on the device while chunk k+1 crosses PCIe; the
 * results are bit-identical to the one-pass ingest.  Environment knobs read at push time:
 * XPCS_NO_PIPELINE, XPCS_PIPELINE_MIN_EVENTS, XPCS_PIPELINE_CHUNKS. */
int xpcs_push_sparse(xpcs_handle h, const int32_t *idx, const int16_t *val,
                     const int64_t *frame_offsets, const double *clock, const double *ticks,
                     int nframes);
/* same, but idx/val/frame_offsets are DEVICE pointers that stay valid (and unmodified)
 * until xpcs_finish_ingest returns; no copy is made.  One call per ingest.  d_idx must be 16-byte aligned,
 * d_val and d_frame_offsets 8-byte aligned (XPCS_E_ARG otherwise): the kernels read four events per load. */
int xpcs_push_sparse_device(xpcs_handle h, const int32_t *d_idx, const int16_t *d_val,
                            const int64_t *d_frame_offsets, int64_t n_events, int nframes);
/* replaces: Imm::NextFrames for uncompressed files + DenseFilter::Apply
 * (filter/dense_filter.cpp:121-210).  frames = [nframes][P] raw int16 (host). */
int xpcs_push_dense(xpcs_handle h, const int16_t *frames, const double *clock,
                    const double *ticks, int nframes);
int xpcs_push_dense_device(xpcs_handle h, const int16_t *d_frames, int nframes);

/* Ends the ingest loop of main.cpp:258-268: builds the pixel-major store (the role of
 * data_structure::SparseData, sparse_data.cpp:59-103) and returns the Filter getters
 * (filter/filter.h:62-96) after the post-scaling of main.cpp:313-343 and :360-378:
 *   pixel_sum[P]        PixelsSum()/F
 *   frame_sum[2F]       FramesSum()
 *   part_total[S]       PartitionsMean()/(pixels_per_sbin*F)
 *   part_partial[floor(F/window)*S]  PartialPartitionsMean()/(pixels_per_sbin*window)
 * all nullable.  With shard_count > 1 and no communicator the sums cover this shard's pixels only; with a
 * communicator (xpcs_comm_init) they are all-reduced and cover the whole detector on every rank. */
int xpcs_finish_ingest(xpcs_handle h, float *pixel_sum, float *frame_sum, float *part_total,
                       float *part_partial);
/* TimestampClock()/TimestampTicks(): [2][raw frames] each (sparse_filter.cpp:134-140) */
int xpcs_get_timestamps(xpcs_handle h, double *clock, double *ticks);
/* replaces the --frameout=N block of main.cpp:276-310: the first `nframes` post-filter frames,
 * out[f * P + pixel] (zeros where a pixel has no event), read back from the pixel-major store.
 * Call between xpcs_finish_ingest and xpcs_multitau.  (With normalize_by_framesum the store
 * already carries the frame-sum scaling, which the reference applies after this dump.) */
int xpcs_get_frames(xpcs_handle h, int nframes, float *out);

/* ---- Correlation ---- */
/* replaces: Corr::multiTau2(SparseData*, float* G2, float* IP, float* IF) (corr.cpp:315-431).
 * G2/IP/IF: host [T][P] tau-major, fully written (zeros for masked / event-less pixels);
 * all three NULL = keep the results on the device only (the normal path without --g2out). */
int xpcs_multitau(xpcs_handle h, float *G2, float *IP, float *IF);
/* The same G2 / IP / IF arrays restricted to `n` listed detector pixels: host [T][n] each (nullable); call
 * after xpcs_multitau.  Pixels that are masked or owned by another shard come back as zeros.  What --g2out
 * gives for a region of interest without moving T*P floats (a 1 Mpixel, 115-delay job is 3 x 483 MB). */
int xpcs_get_correlators(xpcs_handle h, const int32_t *pixels, int n, float *G2, float *IP, float *IF);
/* replaces: Corr::normalizeG2s (corr.cpp:927-1091): g2 and std-error, host (T, Q) each.
 * Equivalent to xpcs_normalize_partials + xpcs_normalize_finish. */
int xpcs_normalize(xpcs_handle h, float *g2, float *stderr_out);

/* Multi-GPU split of xpcs_normalize.  xpcs_normalize_partials leaves, in
 * one device buffer of *count doubles, (a) the normalised g2 row of every static segment
 * this shard owns (zeros elsewhere) and (b) per dynamic bin the sample count, sum and sum
 * of squares of the per-pixel G2/(IP*IF).  Summing the buffers of all shards element-wise
 * (ncclAllReduce / torch.distributed.all_reduce, SUM, float64, in place on *d_partials)
 * yields the single-GPU buffer bit for bit, because every segment is owned by exactly one
 * shard.  xpcs_normalize_finish then produces g2 / stderr from the (reduced) buffer.  With a communicator
 * (xpcs_comm_init) xpcs_normalize_partials performs that all-reduce itself (ncclAllReduce on the handle's
 * stream) and the caller has nothing to exchange. */
int xpcs_normalize_partials(xpcs_handle h, void **d_partials, int64_t *count);
int xpcs_normalize_finish(xpcs_handle h, float *g2, float *stderr_out);

/* ---- online multi-tau: frame streams that do not fit the device ----
 * replaces: the read-everything-then-correlate sequence of main.cpp:263-268 + Corr::multiTau2 (corr.cpp:315-431,
 * whose level binning corr.cpp:349-390 works in place on complete rows) for jobs whose events exceed the device
 * (BASELINE configs[4] at >= 1 % occupancy).  The frames arrive in chunks of chunk_frames = 2^k frames (64..8192);
 * each chunk goes through the Filter stage, is folded into a per-pixel state (integer G2 numerators, per-level
 * totals, the first and last 2*dpl bins of every level) and is then forgotten: device memory is
 * O(pixels * delays + one chunk), independent of the number of frames.  Call sequence:
 *   xpcs_stream_begin(h, chunk_frames)
 *   xpcs_stream_push_sparse[_device](...)   frames in order; every chunk complete except the last of the job
 *   xpcs_stream_finish(h, sums...)          same outputs as xpcs_finish_ingest
 *   xpcs_multitau(h, G2, IP, IF)            hands out the streamed results (no kernel runs)
 *   xpcs_normalize(h, g2, stderr)           unchanged, including the multi-GPU reduction (every rank streams the
 *                                           whole detector's chunks and keeps the pixels of its shard)
 * Results are the exact sums of SURVEY.md A.2/A.3 with the one IEEE division of corr.cpp:420-424: bit-identical to
 * the resident path run without XPCS_COMPAT_STALE_TAIL.  That flag needs the complete rows and is
 * refused (XPCS_E_ARG), as are flat field, stride / averaging, frame-sum normalisation and delays_per_level other
 * than 4 and 8.  xpcs_get_frames and xpcs_twotime need a resident ingest.  A push that fails (e.g. a count beyond
 * the packed word's 4095: XPCS_E_ARG) leaves the stream unusable: xpcs_reset, then start again. */
int xpcs_stream_begin(xpcs_handle h, int chunk_frames);
/* host buffers, layout of xpcs_push_sparse; nframes may span several chunks (cut at multiples of chunk_frames
 * counted from the first frame of the job); a push that ends inside a chunk must be the last one */
int xpcs_stream_push_sparse(xpcs_handle h, const int32_t *idx, const int16_t *val, const int64_t *frame_offsets,
                            const double *clock, const double *ticks, int nframes);
/* device pointers, ONE chunk per call, d_frame_offsets[0] == 0; the buffers may be refilled when the call returns */
int xpcs_stream_push_sparse_device(xpcs_handle h, const int32_t *d_idx, const int16_t *d_val,
                                   const int64_t *d_frame_offsets, int64_t n_events, int nframes);
int xpcs_stream_finish(xpcs_handle h, float *pixel_sum, float *frame_sum, float *part_total, float *part_partial);

/* ---- multi-GPU: one handle per GPU, pixel-sharded ----
 * replaces: the OpenMP pixel loop of Corr::multiTau2 (corr.cpp:329-332) spread over GPUs.  The exchange
 * quantities are the events themselves (below), the per-frame / per-static-bin sums of the Filter stage
 * (sparse_filter.cpp:175,190; corr.cpp:966-1014) and the normalisation partials.  One NCCL communicator rank per
 * handle; NCCL is bound at run time (libnccl.so.2).  Call sequence per rank (one process or one thread per GPU;
 * with a communicator every rank must make the same calls with the same NULL-ness of output arguments, because
 * xpcs_finish_ingest and xpcs_normalize[_partials] contain collectives):
 *   rank 0: xpcs_comm_unique_id(id)  ->  hand the 128 bytes to every rank
 *   all:    xpcs_create(shard_index = rank, shard_count = n)  ->  xpcs_comm_init(h, n, rank, id)
 *   all:    xpcs_push_sparse_slab(h, first_frame_of_my_slab, ...)  ->  xpcs_finish_ingest  ->  xpcs_multitau
 *           ->  xpcs_normalize
 * With a communicator xpcs_finish_ingest returns WHOLE-DETECTOR sums on every rank (frame_sum, part_total,
 * part_partial and pixel_sum are all-reduced; exact for integer counts), xpcs_normalize returns the same g2 /
 * stderr on every rank, bit-identical to a single-GPU run, and normalize_by_framesum works sharded. */
int xpcs_comm_unique_id(void *id128);                                    /* ncclGetUniqueId: 128 bytes out   */
int xpcs_comm_init(xpcs_handle h, int nranks, int rank, const void *id128); /* collective: ncclCommInitRank  */
int xpcs_comm_nccl_version(void);                                        /* 0 = NCCL not loadable            */
/* how the slab exchange of this handle moves the events: 1 = DIRECT, the partition kernel stores every event
 * straight into its owner's list over NVLink (peer access inside a process, CUDA IPC mappings between
 * processes; the transfer is the partition, an all-reduced word is the barrier); 0 = STAGED, per-owner streams
 * and one grouped ncclSend/ncclRecv (ranks on several hosts, no peer access, or XPCS_NO_P2P); -1 = no communicator.
 * Agreed by all ranks at xpcs_comm_init; a rank that cannot map a peer's buffers moves everybody to STAGED. */
int xpcs_comm_transport(xpcs_handle h);
/* Frame-slab ingest: this rank holds raw frames [first_raw_frame, first_raw_frame + nframes) of the WHOLE
 * detector (all pixels) -- e.g. its 1/n of the IMM file, so a job's host->device traffic is the file once,
 * spread over every GPU's PCIe link.  The slabs of ranks 0..n-1 must tile the job's raw frames in rank order.
 * xpcs_finish_ingest then partitions the slab by pixel owner on the device and moves every partition to its
 * owner (grouped ncclSend/ncclRecv over NVLink), after which each rank ingests the events of its own pixels
 * for all frames.  One slab push per ingest.  Arguments as xpcs_push_sparse / xpcs_push_sparse_device
 * (frame_offsets index the slab's own idx/val arrays).  With shard_count == 1 these are plain pushes. */
int xpcs_push_sparse_slab(xpcs_handle h, int first_raw_frame, const int32_t *idx, const int16_t *val,
                          const int64_t *frame_offsets, const double *clock, const double *ticks, int nframes);
int xpcs_push_sparse_slab_device(xpcs_handle h, int first_raw_frame, const int32_t *d_idx, const int16_t *d_val,
                                 const int64_t *d_frame_offsets, int64_t n_events, int nframes);

/* replaces: Corr::twotime -> twotimeQBinThreading (corr.cpp:562-572, :781-924) including
 * Smoothing (:433-560, :1166-1305), for ONE dynamic bin `qbin`:
 *   C[F*F] row-major upper triangle (lower = 0), g2full[F], g2partials[wsize*partials]
 *   ([d][w]), sg[rows][F] (or [rows][1] when average) -- all host, nullable.
 * smoothing_method: 0 = none, 1 = symmetric (one sg row: the bin's mean intensity per frame,
 * ComputeSGSymmetric :1166-1226), 2 = StaticMap (one sg row per static partition of the bin, in the
 * order of Configuration::getBinMaps(); every pixel is divided by the sg of its own static partition,
 * SmoothingStaticMap :433-494, ComputeSGStaticMap :1228-1305; sg must hold n_static * F floats);
 * smoothing_average: the "Average" filter.  *sg_rows (nullable) receives the number of sg rows.
 * twotimeFrameThreading (corr.cpp:574-779, `corr --frame_threading`) computes the same C in another
 * summation order on the CPU; there is one contraction here (tensor cores), equal to both within 1e-5. */
int xpcs_twotime_sg(xpcs_handle h, int qbin, int wsize, int smoothing_method, int smoothing_average,
                    float *C, float *g2full, float *g2partials, float *sg, int *sg_rows);
/* the same without the row count (symmetric smoothing: sg[F] or sg[1]) */
int xpcs_twotime(xpcs_handle h, int qbin, int wsize, int smoothing_method, int smoothing_average,
                 float *C, float *g2full, float *g2partials, float *sg);

/* ---- pinned host memory ----
 * Page-locked host buffers (cudaHostAlloc) for callers that do not link the CUDA runtime themselves: pushes from
 * and result copies into such buffers are asynchronous and run at the full PCIe rate (pageable memory is staged
 * by the driver at a fraction of it).  Returns NULL on failure. */
void *xpcs_host_alloc(size_t bytes);
void xpcs_host_free(void *p);

/* ---- measurement hooks (bench.py, profiles/) ---- */
/* record a CUDA event pair around every kernel launch of this handle */
int xpcs_kernel_timing(xpcs_handle h, int enable);
/* kernels launched by this handle since the last xpcs_kernel_report_reset */
int64_t xpcs_launch_count(xpcs_handle h);
/* per-kernel totals since the last reset: writes up to cap entries, returns the number of
 * distinct kernels.  names[i] points into handle-owned storage. */
int xpcs_kernel_report(xpcs_handle h, const char **names, double *total_ms, int64_t *launches,
                       int cap);
int xpcs_kernel_report_reset(xpcs_handle h);

/* slices (32 pixel rows) the warp-per-row multi-tau kernel left to the lane-per-row kernel in
 * the last xpcs_multitau (rows beyond its shared-memory budget or with counts summing to 2^16
 * or more); -1 when the warp kernel did not run (float values, unusual delays-per-level). */
int64_t xpcs_multitau_fallback_slices(xpcs_handle h);

/* library/ABI version, and the compute capability it was compiled for (100 = sm_100a) */
int xpcs_abi_version(void);
int xpcs_compiled_arch(void);

#ifdef __cplusplus
}
#endif
#endif /* XPCS_B200_H */
