// internal.h -- handle layout and kernel launchers shared by the .cu files of libxpcs_b200.
// Not part of the C-ABI (include/xpcs_b200.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/xpcs_b200.h"

namespace xpcs {

constexpr int kSlice = 32;          // pixel rows per slice = lanes per warp
constexpr int kCountBits = 12;      // packed word at level l: key << (12+l) | count
constexpr int kMaxLevels = 24;
constexpr int kEvPerBlock = 4096;   // events per CTA in the frame-major passes
constexpr int kIngestThreads = 256;

enum ValueKind { kPacked = 0, kFloat = 1 };

// Device-side description of the correlation job of one handle.
struct Sched {
    int n_levels;               // max_level + 1
    int frames;                 // F
    int dpl;
    int first[kMaxLevels];      // index of the level's first delay in the schedule
    int count[kMaxLevels];      // delays emitted at that level
    int lo[kMaxLevels];         // level-local delay tau' of the level's first delay
};

struct KernelStat {
    double ms = 0.0;
    int64_t launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

// Arguments of the multi-tau kernels (multitau.cu: one lane per row; multitau_warp.cu: one
// warp per row).
struct MtArgs {
    void *store;
    const int64_t *slice_base;
    const int *slice_len;
    const int *row_len;
    float *G2, *IP, *IF;
    int R_pad, n_slices, smem_len, hi, compat;
    const unsigned char *only_flagged;  // lane-per-row kernel: process only slices flagged here (nullptr = all)
    Sched sched;
};

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;  // capacity in elements
};

// Pipelined sparse ingest: the finished pixel-major store of one chunk of frames.
constexpr int kMaxChunks = 16;
struct ChunkStore {
    DevBuf<uint32_t> store;
    DevBuf<int64_t> slice_base;
    DevBuf<int> row_len;
};

}  // namespace xpcs

struct xpcs_handle_s {
    XpcsParams prm{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;

    // ---- schedule (host) ----
    std::vector<int> sched_level, sched_tau;
    int T = 0, max_level = 0;
    xpcs::Sched sched{};

    // ---- partition maps (host) ----
    int P = 0, S = 0, Q = 0;
    int nseg_total = 0;                 // surviving (dq, sq) entries of the whole detector
    std::vector<int> seg_dq, seg_sq;    // [nseg_total]
    std::vector<int> seg_pixels_n;      // [nseg_total] pixels per segment
    int seg_first = 0, seg_last = 0;    // this shard owns global segments [seg_first, seg_last)
    int R = 0, R_pad = 0;               // rows of this shard, padded to kSlice
    int R_total = 0;
    int n_slices = 0;
    std::vector<int> pixel_of_row;      // [R]
    std::vector<int> lseg_row_start;    // [nseg_local + 1] row offsets of the owned segments
    std::vector<int> pixels_per_sbin;   // [S]
    std::vector<double> flat_host;      // [P] (all ones when disabled)
    bool flat_is_one = true;

    // ---- maps (device) ----
    xpcs::DevBuf<int> d_row_of_pixel;     // [P]   -1 = masked or foreign shard
    xpcs::DevBuf<int> d_pixel_of_row;     // [R_pad]
    xpcs::DevBuf<int> d_sbin_of_row;      // [R_pad] sq-1
    xpcs::DevBuf<double> d_flat;          // [P]
    xpcs::DevBuf<int> d_lseg_row_start;   // [nseg_local+1]
    xpcs::DevBuf<int> d_seg_dq_all;       // [nseg_total] dynamic bin of every segment (all shards)
    xpcs::DevBuf<int> d_seg_npix_all;     // [nseg_total] pixels per segment (all shards)

    // ---- dark image ----
    bool have_dark = false;
    xpcs::DevBuf<double> d_dark_avg, d_dark_std;
    xpcs::DevBuf<int16_t> d_dense_bound;        // [P] dense filter: a sample survives iff raw > bound
    xpcs::DevBuf<unsigned char> d_dense_every;  // [P/8] group holds a pixel that lets every raw value through
    bool dense_bounds_ready = false;
    xpcs::DevBuf<unsigned char> d_dense_args;   // launch-invariant dense filter arguments in global memory
    bool dense_args_ready = false;
    // push_dense: double-buffered device staging, H2D on its own stream
    xpcs::DevBuf<int16_t> d_dense_stage[2];
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_filtered[2] = {nullptr, nullptr};

    // ---- frame-major event buffers ----
    bool dense_source = false;
    bool external_events = false;         // push_sparse_device: buffers are the caller's
    xpcs::DevBuf<int32_t> d_idx;
    xpcs::DevBuf<int16_t> d_val;
    xpcs::DevBuf<int32_t> d_evt;          // dense source: explicit output-frame id per event
    xpcs::DevBuf<float> d_valf;           // dense source: final float value per event
    const int32_t *ev_idx = nullptr;      // views actually used by the kernels
    const int16_t *ev_val = nullptr;
    const int64_t *ev_off = nullptr;      // device frame offsets [raw_frames + 1]
    xpcs::DevBuf<int64_t> d_frame_off;
    std::vector<int64_t> frame_off_host;  // [raw_frames + 1]
    int64_t E = 0;                        // events pushed
    int raw_frames = 0;
    std::vector<double> ts_clock, ts_ticks;
    xpcs::DevBuf<unsigned long long> d_dense_counter;
    // pinned staging for small pushes
    void *stage = nullptr;
    size_t stage_bytes = 0;

    // ---- pixel-major store (slices of 32 rows, column-interleaved words) ----
    int kind = xpcs::kPacked;
    bool ingest_done = false;
    xpcs::DevBuf<int> d_row_count;        // [R_pad] histogram, consumed by the scatter
    xpcs::DevBuf<int> d_row_len;          // [R_pad] events per row
    xpcs::DevBuf<int> d_slice_len;        // [n_slices]
    xpcs::DevBuf<int64_t> d_slice_base;   // [n_slices + 1] in words
    xpcs::DevBuf<int> d_slice_cur;        // [n_slices] events of the slice, counted down by the stream scatter
    xpcs::DevBuf<int64_t> d_slice_rec;    // [n_slices + 1] first record of the slice in d_rec
    xpcs::DevBuf<unsigned long long> d_slice_end;  // [n_slices] end of the slice's stream, counted down by the stream scatter
    xpcs::DevBuf<unsigned long long> d_rec;  // [E] slice-ordered event records (lane | word)
    xpcs::DevBuf<int> d_block_first;      // first raw frame of every event block
    xpcs::DevBuf<uint32_t> d_store;       // words (uint32 packed, or pairs for float values)
    int64_t store_words = 0;
    int64_t events_stored = 0;
    int max_row = 0;
    int max_count = 0;                    // largest merged photon count of the packed store (two-time: fp16 is exact up to 2048)
    xpcs::DevBuf<long long> d_summary;    // small device scratch for host-visible scalars
    xpcs::DevBuf<double> d_frame_acc;     // [F] per-output-frame sums (exact for counts)
    xpcs::DevBuf<double> d_row_sum;       // [R_pad]
    xpcs::DevBuf<double> d_part_total;    // [S]
    xpcs::DevBuf<double> d_part_partial;  // [windows * S]
    xpcs::DevBuf<float> d_frame_scale;    // [F] normalize_by_framesum divisors
    // pipelined sparse ingest (xpcs_push_sparse with host buffers, integer counts): every chunk of
    // frames is ingested on the handle's stream while the next one crosses PCIe on the copy stream;
    // xpcs_finish_ingest concatenates the chunk stores
    bool pipe_on = false, pipe_broken = false;
    int pipe_chunks = 0;                  // chunk stores filled in this ingest
    int64_t pipe_events = 0;              // events the chunk ingests stored (before merging duplicates)
    xpcs::ChunkStore chunk[xpcs::kMaxChunks];
    cudaEvent_t ev_chunk[xpcs::kMaxChunks] = {};
    int64_t frame_off_uploaded = 0;       // entries of frame_off_host already in d_frame_off
    std::vector<float> frame_sum_host;    // [2F]

    // ---- online multi-tau (multitau_stream.cu): frames arrive in chunks of 2^stream_k frames, a per-row state is all
    // that survives a chunk
    bool stream_on = false, stream_done = false;
    int stream_k = 0;
    int stream_chunks = 0;                // complete or final chunks consumed so far
    bool stream_short_seen = false;       // a chunk shorter than 2^stream_k came in: it has to be the last one
    xpcs::DevBuf<uint32_t> d_stream_state;  // [R_pad][state words]
    // host pushes of several chunks: chunk k + 1 crosses PCIe into the second set of event buffers while chunk k
    // (in d_idx / d_val / d_frame_off) is ingested and folded into the state; the sets swap after every chunk
    xpcs::DevBuf<int32_t> d_st_idx2;
    xpcs::DevBuf<int16_t> d_st_val2;
    xpcs::DevBuf<int64_t> d_st_off2;
    cudaEvent_t ev_st_copy[2] = {nullptr, nullptr};

    // ---- results ----
    bool multitau_done = false;
    bool mt_warp_ran = false;             // last multi-tau used the warp-per-row kernel
    bool rows_consumed = false;
    xpcs::DevBuf<float> d_G2, d_IP, d_IF;   // [T][R_pad] row-permuted, tau-major
    xpcs::DevBuf<unsigned char> d_mt_fallback;  // [n_slices] slices the warp-per-row kernel leaves to the lane-per-row one
    xpcs::DevBuf<double> d_partials;        // see xpcs_normalize_partials
    int64_t partials_count = 0;
    bool partials_done = false;
    xpcs::DevBuf<float> d_scratch;          // staging for host-layout outputs

    // ---- two-time (one dynamic partition at a time) ----
    xpcs::DevBuf<unsigned short> d_tt_hi, d_tt_lo;  // fp16 operands Xt[F][npad] (value = hi + lo)
    xpcs::DevBuf<float> d_tt_C;                     // [F][F]
    xpcs::DevBuf<float> d_tt_sg, d_tt_out;
    xpcs::DevBuf<unsigned int> d_tt_sgint;
    xpcs::DevBuf<double> d_tt_diag;
    xpcs::DevBuf<int> d_tt_tiles;                   // (row tile, column tile) pairs of the GEMM grid

    // ---- multi-GPU (comm.cu) ----
    void *comm = nullptr;                 // ncclComm_t
    int comm_nranks = 1, comm_rank = 0;
    std::vector<unsigned char> owner_of_pixel;     // [P] shard that owns the pixel, 255 = masked (plan_maps)
    xpcs::DevBuf<unsigned char> d_owner_of_pixel;  // uploaded at xpcs_comm_init
    // frame-slab ingest: this handle was given raw frames [slab_first, slab_first + slab_frames) of the WHOLE
    // detector; xpcs_finish_ingest redistributes the events to the pixel owners
    bool slab_mode = false;
    int slab_first = 0, slab_frames = 0;
    int64_t slab_events = 0;
    const int32_t *slab_idx = nullptr;    // device views (the caller's buffers or the d_slab_* copies)
    const int16_t *slab_val = nullptr;
    const int64_t *slab_off = nullptr;
    xpcs::DevBuf<int32_t> d_slab_idx;
    xpcs::DevBuf<int16_t> d_slab_val;
    xpcs::DevBuf<int64_t> d_slab_off;
    xpcs::DevBuf<int64_t> d_dm_off;       // [nranks][slab_frames + 1] per-destination event offsets of the slab's frames
    xpcs::DevBuf<int64_t> d_dm_meta;      // [nranks + 2] first frame, frames, events per destination
    xpcs::DevBuf<int64_t> d_dm_all;       // [nranks][nranks + 2] the same of every rank
    xpcs::DevBuf<int32_t> d_send_idx;
    xpcs::DevBuf<int16_t> d_send_val;
    xpcs::DevBuf<int64_t> d_recv_off;     // [raw frames + nranks] offsets as received, one stream per source rank
    // direct NVLink stores (comm.cu): the partition kernel writes every event straight into its owner's list
    bool p2p_enabled = false;             // all ranks on one host, peer access possible, not disabled by XPCS_NO_P2P
    bool p2p_mapped = false;              // peer pointers below are current
    long long p2p_gen = 0;                // bumped whenever this rank's receive buffers move
    std::vector<long long> peer_pid, peer_gen_seen;
    std::vector<int32_t *> peer_idx;      // [nranks] the ranks' d_idx / d_val / d_recv_off as seen from this device
    std::vector<int16_t *> peer_val;
    std::vector<int64_t *> peer_off;
    std::vector<void *> peer_opened;      // IPC mappings to close
    xpcs::DevBuf<int64_t> d_p2p_xchg;     // mapping records, mine + everyone's
    bool frame_acc_reduced = false;       // the per-frame sums already cover every shard
    bool part_sums_reduced = false;       // so do the per-static-bin sums
    std::vector<int> slab_first_of_rank, slab_frames_of_rank;  // filled by the exchange

    // ---- measurement ----
    bool timing = false;
    int64_t launches = 0;
    std::map<std::string, xpcs::KernelStat> stats;
};

namespace xpcs {

// RAII bracket around one kernel launch: counts it and, when timing is on, records a CUDA
// event pair on the handle's stream.
struct LaunchScope {
    xpcs_handle_s *h;
    const char *name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    // own = false: a library kernel (NCCL) -- timed under its name, not counted as one of ours
    LaunchScope(xpcs_handle_s *h_, const char *name_, bool own = true);
    ~LaunchScope();
};

int fail(xpcs_handle_s *h, int code, const char *fmt, ...);
int check_cuda(xpcs_handle_s *h, cudaError_t e, const char *what);

template <typename T>
int ensure(xpcs_handle_s *h, DevBuf<T> &b, size_t n, const char *what)
{
    if (b.n >= n && b.p) return XPCS_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.n = 0;
    size_t want = n ? n : 1;
    cudaError_t e = cudaMalloc((void **)&b.p, want * sizeof(T));
    if (e != cudaSuccess) return check_cuda(h, e, what);
    b.n = want;
    return XPCS_OK;
}

template <typename T>
void release(DevBuf<T> &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.n = 0;
}

// ---- cross-GPU exchange (comm.cu): NCCL bound at run time, one communicator per handle ----
bool comm_active(const xpcs_handle_s *h);                            // communicator with more than one rank
int comm_allreduce_f64(xpcs_handle_s *h, double *d_buf, size_t n);   // in place, SUM, on the handle's stream
int comm_allreduce_f32(xpcs_handle_s *h, float *d_buf, size_t n);
int comm_allreduce_f64_group(xpcs_handle_s *h, double *const *bufs, const size_t *counts, int n);
int comm_exchange_slab(xpcs_handle_s *h);                            // frame slabs -> pixel shards (xpcs_push_sparse_slab*)
void comm_destroy(xpcs_handle_s *h);

// ---- launchers (ingest.cu) ----
int launch_ingest(xpcs_handle_s *h);             // histogram -> slices -> scatter -> finalize
int launch_ingest_chunk(xpcs_handle_s *h, int f0, int f1);  // same for raw frames [f0, f1) into the next chunk store; 1 = not representable
int launch_ingest_concat(xpcs_handle_s *h);      // chunk stores -> the store
// stream mode: the events d_idx/d_val[0, nev) with local frame offsets d_frame_off[0..nframes] are the raw frames
// [frame_base, frame_base + nframes); their store replaces chunk[0], the Filter sums accumulate unless `first`
int launch_ingest_stream_chunk(xpcs_handle_s *h, int frame_base, int nframes, int64_t nev, bool first);
int launch_dark(xpcs_handle_s *h, const int16_t *d_frames, int n);
int launch_dense_filter(xpcs_handle_s *h, const int16_t *d_frames, int first_raw, int nframes);
// ---- launchers (multitau.cu) ----
int launch_multitau(xpcs_handle_s *h);
int launch_unpermute(xpcs_handle_s *h, const float *d_src, float *d_dst);  // [T][R_pad] -> [T][P]
// ---- launchers (multitau_warp.cu) ----
bool multitau_warp_eligible(const xpcs_handle_s *h);
int launch_multitau_warp(xpcs_handle_s *h, MtArgs &a);   // fills h->d_mt_fallback
// ---- launchers (multitau_slice.cu) ----
bool multitau_slice_eligible(const xpcs_handle_s *h);
int launch_multitau_slice(xpcs_handle_s *h, MtArgs &a);  // fills h->d_mt_fallback
// ---- launchers (multitau_warpf.cu) ----
bool multitau_warpf_eligible(const xpcs_handle_s *h);
int launch_multitau_warpf(xpcs_handle_s *h, MtArgs &a, bool flagged_only = false);  // fills / updates h->d_mt_fallback
// lane = row kernel for float rows whose slices fit a shared-memory tile (multitau_slicef.cu)
bool multitau_slicef_eligible(const xpcs_handle_s *h);
int launch_multitau_slicef(xpcs_handle_s *h, MtArgs &a);  // fills h->d_mt_fallback
// ---- launchers (multitau_stream.cu) ----
int stream_check(xpcs_handle_s *h, int chunk_frames);  // can this job be streamed with that chunk length?
int launch_stream_begin(xpcs_handle_s *h);             // allocates and clears the per-row state
int launch_stream_chunk(xpcs_handle_s *h, int c, bool empty);  // chunk c from chunk[0] (empty: no event at all)
int launch_stream_finish(xpcs_handle_s *h);            // state -> d_G2 / d_IP / d_IF
// ---- launchers (normalize.cu) ----
int launch_normalize_partials(xpcs_handle_s *h);
int launch_normalize_finish(xpcs_handle_s *h, float *d_g2, float *d_se);
// ---- launchers (twotime.cu) ----
int launch_twotime(xpcs_handle_s *h, int qbin, int wsize, int method, int average, float *C,
                   float *g2full, float *g2partials, float *sg, int *sg_rows_out);

}  // namespace xpcs
