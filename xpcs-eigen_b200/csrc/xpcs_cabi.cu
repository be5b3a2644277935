// xpcs_cabi.cu -- the C-ABI of include/xpcs_b200.h: handle lifetime, partition maps,
// delay schedule, host<->device plumbing around the kernels of ingest.cu, multitau.cu,
// normalize.cu and twotime.cu.  No CPU compute path exists here: every stage runs as a CUDA
// kernel and every entry point fails when no device is usable.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <numeric>

#include "internal.h"

using namespace xpcs;

static thread_local std::string g_create_error;

namespace xpcs {

int fail(xpcs_handle_s *h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

int check_cuda(xpcs_handle_s *h, cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return XPCS_OK;
    return fail(h, XPCS_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

LaunchScope::LaunchScope(xpcs_handle_s *h_, const char *name_, bool own) : h(h_), name(name_)
{
    if (own) h->launches++;
    if (h->timing) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, h->stream);
    }
}

LaunchScope::~LaunchScope()
{
    KernelStat &st = h->stats[name];
    st.launches++;
    if (h->timing && e0) {
        cudaEventRecord(e1, h->stream);
        st.pending.emplace_back(e0, e1);
    }
}

}  // namespace xpcs

// ---------------------------------------------------------------------------------------
// schedule: Corr::calculateLevelMax / Corr::delaysPerLevel (reference corr.cpp:1133-1160)
// ---------------------------------------------------------------------------------------
extern "C" int xpcs_level_max(int frames, int dpl)
{
    if (dpl <= 0 || frames < dpl * 2) return 0;
    return (int)(floor(log2((double)frames) - log2(1.0 + 1.0 / (double)dpl)) - log2((double)dpl));
}

extern "C" int xpcs_delay_schedule(int frames, int dpl, int32_t *level, int32_t *tau, int cap)
{
    if (dpl <= 0 || frames <= 0) return 0;
    const int top = xpcs_level_max(frames, dpl);
    int n = 0;
    long long reached = 0;
    for (int lv = 0; lv <= top; lv++) {
        const long long step = 1LL << lv;
        const int quota = lv == 0 ? 2 * dpl : dpl;
        for (int j = 0; j < quota; j++) {
            if (reached + 2 * step > frames) break;
            reached += step;
            if (n < cap) {
                if (level) level[n] = lv;
                if (tau) tau[n] = (int32_t)reached;
            }
            n++;
        }
    }
    return n;
}

static int build_schedule(xpcs_handle_s *h)
{
    const int F = h->prm.frames, dpl = h->prm.delays_per_level;
    h->max_level = xpcs_level_max(F, dpl);
    if (h->max_level + 1 > kMaxLevels) return fail(h, XPCS_E_ARG, "too many levels (%d)", h->max_level + 1);
    const int T = xpcs_delay_schedule(F, dpl, nullptr, nullptr, 0);
    h->T = T;
    h->sched_level.assign(T, 0);
    h->sched_tau.assign(T, 0);
    xpcs_delay_schedule(F, dpl, h->sched_level.data(), h->sched_tau.data(), T);
    Sched &s = h->sched;
    memset(&s, 0, sizeof(s));
    s.n_levels = h->max_level + 1;
    s.frames = F;
    s.dpl = dpl;
    for (int l = 0; l < s.n_levels; l++) { s.first[l] = 0; s.count[l] = 0; s.lo[l] = 1; }
    for (int i = 0; i < T; i++) {
        const int l = h->sched_level[i];
        const int local = h->sched_tau[i] >> l;  // tau / 2^level, exact (corr.cpp:392-393)
        if ((local << l) != h->sched_tau[i]) return fail(h, XPCS_E_ARG, "schedule: tau not a multiple of 2^level");
        if (s.count[l] == 0) { s.first[l] = i; s.lo[l] = local; }
        else if (local != s.lo[l] + s.count[l]) return fail(h, XPCS_E_ARG, "schedule: delays of a level not consecutive");
        s.count[l]++;
        if (local > 2 * dpl) return fail(h, XPCS_E_ARG, "schedule: level-local delay beyond 2*dpl");
    }
    return XPCS_OK;
}

// ---------------------------------------------------------------------------------------
// partition maps: Configuration::BuildQMap (reference configuration.cpp:244-381)
// ---------------------------------------------------------------------------------------
struct Cell {
    int dq, sq;
    int64_t start;  // into the sorted valid-pixel list
    int n;
    bool kept;
};

// host-only part: partition maps, segments, shard cut, row order (no CUDA call)
static int plan_maps(xpcs_handle_s *h)
{
    const XpcsParams &p = h->prm;
    const int P = h->P;
    std::vector<int> valid;
    valid.reserve(P);
    int S = 0, Q = 0;
    for (int i = 0; i < P; i++) {
        if (p.dqmap[i] < 1 || p.sqmap[i] < 1) continue;  // configuration.cpp:256
        valid.push_back(i);
        Q = std::max(Q, p.dqmap[i]);
        S = std::max(S, p.sqmap[i]);
    }
    h->S = S;
    h->Q = Q;
    h->R_total = (int)valid.size();
    h->pixels_per_sbin.assign(S, 0);
    for (int i : valid) h->pixels_per_sbin[p.sqmap[i] - 1]++;  // configuration.cpp:347-356

    // (dq, sq, pixel) order == iteration order of the reference's nested std::map
    std::stable_sort(valid.begin(), valid.end(), [&](int a, int b) {
        if (p.dqmap[a] != p.dqmap[b]) return p.dqmap[a] < p.dqmap[b];
        if (p.sqmap[a] != p.sqmap[b]) return p.sqmap[a] < p.sqmap[b];
        return a < b;
    });
    std::vector<Cell> cells;
    for (size_t i = 0; i < valid.size();) {
        size_t j = i;
        while (j < valid.size() && p.dqmap[valid[j]] == p.dqmap[valid[i]] && p.sqmap[valid[j]] == p.sqmap[valid[i]]) j++;
        cells.push_back(Cell{p.dqmap[valid[i]], p.sqmap[valid[i]], (int64_t)i, (int)(j - i), true});
        i = j;
    }
    // A static bin listed under several dynamic bins survives only under the one with most
    // pixels (first such in ascending dq); the others' pixels are lost (configuration.cpp:311-345,
    // SURVEY.md A.8-1).
    {
        std::map<int, int> best;  // sq -> cell index
        std::map<int, int> owners;
        for (size_t c = 0; c < cells.size(); c++) {
            owners[cells[c].sq]++;
            auto it = best.find(cells[c].sq);
            if (it == best.end() || cells[c].n > cells[it->second].n) best[cells[c].sq] = (int)c;
        }
        for (size_t c = 0; c < cells.size(); c++)
            if (owners[cells[c].sq] > 1 && best[cells[c].sq] != (int)c) cells[c].kept = false;
    }
    h->seg_dq.clear();
    h->seg_sq.clear();
    h->seg_pixels_n.clear();
    std::vector<const Cell *> kept;
    int64_t kept_pixels = 0;
    for (const Cell &c : cells)
        if (c.kept) {
            kept.push_back(&c);
            h->seg_dq.push_back(c.dq);
            h->seg_sq.push_back(c.sq);
            h->seg_pixels_n.push_back(c.n);
            kept_pixels += c.n;
        }
    h->nseg_total = (int)kept.size();

    // shard = contiguous segment range balanced by pixel count
    const int K = p.shard_count, k = p.shard_index;
    std::vector<int> cut(K + 1, h->nseg_total);
    cut[0] = 0;
    {
        int64_t run = 0;
        int next = 1;
        for (int s = 0; s < h->nseg_total && next < K; s++) {
            run += kept[s]->n;
            while (next < K && run * K >= kept_pixels * next) cut[next++] = s + 1;
        }
        for (int i = 1; i <= K; i++) cut[i] = std::max(cut[i], cut[i - 1]);
        cut[K] = h->nseg_total;
    }
    h->seg_first = cut[k];
    h->seg_last = cut[k + 1];
    // which shard owns which pixel (the frame-slab exchange of comm.cu partitions by it); 255 = masked
    h->owner_of_pixel.assign((size_t)P, 255);
    if (K <= 254) {
        for (int sh = 0; sh < K; sh++)
            for (int s = cut[sh]; s < cut[sh + 1]; s++)
                for (int i = 0; i < kept[s]->n; i++) h->owner_of_pixel[valid[kept[s]->start + i]] = (unsigned char)sh;
        for (const Cell &c : cells)  // pixels lost by the duplicate removal: trailing rows of the last shard
            if (!c.kept)
                for (int i = 0; i < c.n; i++) h->owner_of_pixel[valid[c.start + i]] = (unsigned char)(K - 1);
    }

    h->pixel_of_row.clear();
    h->lseg_row_start.assign(1, 0);
    for (int s = h->seg_first; s < h->seg_last; s++) {
        for (int i = 0; i < kept[s]->n; i++) h->pixel_of_row.push_back(valid[kept[s]->start + i]);
        h->lseg_row_start.push_back((int)h->pixel_of_row.size());
    }
    if (k == K - 1) {  // pixels lost by the duplicate removal still get correlated (they are
                       // in the mask), they just belong to no partition: last shard, trailing rows
        std::vector<int> orphans;
        for (const Cell &c : cells)
            if (!c.kept)
                for (int i = 0; i < c.n; i++) orphans.push_back(valid[c.start + i]);
        std::sort(orphans.begin(), orphans.end());
        for (int px : orphans) h->pixel_of_row.push_back(px);
    }
    h->R = (int)h->pixel_of_row.size();
    h->R_pad = (h->R + kSlice - 1) / kSlice * kSlice;
    if (h->R_pad == 0) h->R_pad = kSlice;
    h->n_slices = h->R_pad / kSlice;
    return XPCS_OK;
}

static int build_maps(xpcs_handle_s *h)
{
    int rc0 = plan_maps(h);
    if (rc0) return rc0;
    const XpcsParams &p = h->prm;
    const int P = h->P;
    // device copies
    std::vector<int> row_of_pixel(P, -1), sbin_of_row(h->R_pad, -1), pix_of_row(h->R_pad, 0);
    for (int r = 0; r < h->R; r++) {
        row_of_pixel[h->pixel_of_row[r]] = r;
        sbin_of_row[r] = p.sqmap[h->pixel_of_row[r]] - 1;
        pix_of_row[r] = h->pixel_of_row[r];
    }
    int rc;
    if ((rc = ensure(h, h->d_row_of_pixel, (size_t)P, "row_of_pixel"))) return rc;
    if ((rc = ensure(h, h->d_pixel_of_row, (size_t)h->R_pad, "pixel_of_row"))) return rc;
    if ((rc = ensure(h, h->d_sbin_of_row, (size_t)h->R_pad, "sbin_of_row"))) return rc;
    if ((rc = ensure(h, h->d_flat, (size_t)P, "flatfield"))) return rc;
    if ((rc = ensure(h, h->d_lseg_row_start, h->lseg_row_start.size(), "segment rows"))) return rc;
    if ((rc = ensure(h, h->d_seg_dq_all, (size_t)std::max(1, h->nseg_total), "segment dq"))) return rc;
    if ((rc = ensure(h, h->d_seg_npix_all, (size_t)std::max(1, h->nseg_total), "segment sizes"))) return rc;
    if ((rc = ensure(h, h->d_row_count, (size_t)h->R_pad, "row histogram"))) return rc;
    if ((rc = ensure(h, h->d_row_len, (size_t)h->R_pad, "row lengths"))) return rc;
    if ((rc = ensure(h, h->d_slice_len, (size_t)h->n_slices, "slice lengths"))) return rc;
    if ((rc = ensure(h, h->d_slice_base, (size_t)h->n_slices + 1, "slice offsets"))) return rc;
    if ((rc = ensure(h, h->d_slice_cur, (size_t)h->n_slices + 1, "slice cursors"))) return rc;
    if ((rc = ensure(h, h->d_slice_rec, (size_t)h->n_slices + 1, "slice record offsets"))) return rc;
    if ((rc = ensure(h, h->d_slice_end, (size_t)h->n_slices + 1, "slice stream cursors"))) return rc;
    cudaMemcpy(h->d_row_of_pixel.p, row_of_pixel.data(), sizeof(int) * P, cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_pixel_of_row.p, pix_of_row.data(), sizeof(int) * h->R_pad, cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_sbin_of_row.p, sbin_of_row.data(), sizeof(int) * h->R_pad, cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_flat.p, h->flat_host.data(), sizeof(double) * P, cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_lseg_row_start.p, h->lseg_row_start.data(), sizeof(int) * h->lseg_row_start.size(),
               cudaMemcpyHostToDevice);
    if (h->nseg_total > 0) {
        cudaMemcpy(h->d_seg_dq_all.p, h->seg_dq.data(), sizeof(int) * h->nseg_total, cudaMemcpyHostToDevice);
        cudaMemcpy(h->d_seg_npix_all.p, h->seg_pixels_n.data(), sizeof(int) * h->nseg_total, cudaMemcpyHostToDevice);
    }
    cudaMemset(h->d_row_count.p, 0, sizeof(int) * h->R_pad);
    cudaMemset(h->d_row_len.p, 0, sizeof(int) * h->R_pad);
    return check_cuda(h, cudaGetLastError(), "map upload");
}

// ---------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------
extern "C" int xpcs_abi_version(void) { return 1; }

extern "C" int xpcs_compiled_arch(void)
{
#ifdef XPCS_ARCH
    return XPCS_ARCH;
#else
    return 100;
#endif
}

extern "C" const char *xpcs_last_error(xpcs_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static void reset_ingest(xpcs_handle_s *h)
{
    h->E = 0;
    h->raw_frames = 0;
    h->frame_off_host.assign(1, 0);
    h->ts_clock.clear();
    h->ts_ticks.clear();
    h->ingest_done = false;
    h->multitau_done = false;
    h->partials_done = false;
    h->rows_consumed = false;
    h->dense_source = false;
    h->dense_args_ready = false;
    h->external_events = false;
    h->ev_idx = nullptr;
    h->ev_val = nullptr;
    h->ev_off = nullptr;
    h->events_stored = 0;
    h->store_words = 0;
    h->max_row = 0;
    h->max_count = 0;
    h->pipe_on = false;
    h->pipe_broken = false;
    h->pipe_chunks = 0;
    h->frame_off_uploaded = 0;
    h->slab_mode = false;
    h->slab_first = 0;
    h->slab_frames = 0;
    h->slab_events = 0;
    h->slab_idx = nullptr;
    h->slab_val = nullptr;
    h->slab_off = nullptr;
    h->frame_acc_reduced = false;
    h->part_sums_reduced = false;
    h->stream_on = false;
    h->stream_done = false;
    h->stream_chunks = 0;
    h->stream_short_seen = false;
}

static int check_params(const XpcsParams *prm)
{
    if (!prm || prm->struct_size != (int32_t)sizeof(XpcsParams))
        return fail(nullptr, XPCS_E_ARG, "XpcsParams.struct_size mismatch (caller %d, library %d)",
                    prm ? prm->struct_size : -1, (int)sizeof(XpcsParams));
    if (prm->width <= 0 || prm->height <= 0 || prm->frames <= 0 || prm->delays_per_level <= 0 ||
        prm->stride_frames <= 0 || prm->avg_frames <= 0 || prm->static_window <= 0 || !prm->dqmap ||
        !prm->sqmap || prm->shard_count <= 0 || prm->shard_index < 0 || prm->shard_index >= prm->shard_count)
        return fail(nullptr, XPCS_E_ARG, "invalid XpcsParams (dimensions, frames, dpl, stride/avg, window, maps or shard)");
    if ((int64_t)prm->width * prm->height > 0x7fffffffLL) return fail(nullptr, XPCS_E_ARG, "detector too large");
    if (prm->shard_count > 254) return fail(nullptr, XPCS_E_ARG, "at most 254 shards");
    return XPCS_OK;
}

extern "C" int xpcs_plan_shard(const XpcsParams *prm, XpcsShardPlan *plan, int32_t *row_pixels, int64_t cap)
{
    int rc = check_params(prm);
    if (rc) return rc;
    if (!plan) return fail(nullptr, XPCS_E_ARG, "plan is NULL");
    xpcs_handle_s tmp;
    tmp.prm = *prm;
    tmp.P = prm->width * prm->height;
    if ((rc = build_schedule(&tmp)) || (rc = plan_maps(&tmp))) {
        g_create_error = tmp.err;
        return rc;
    }
    plan->n_static = tmp.S;
    plan->n_dynamic = tmp.Q;
    plan->n_segments = tmp.nseg_total;
    plan->seg_first = tmp.seg_first;
    plan->seg_last = tmp.seg_last;
    plan->n_rows = tmp.R;
    plan->n_rows_total = tmp.R_total;
    plan->n_delays = tmp.T;
    if (row_pixels)
        for (int r = 0; r < tmp.R && r < cap; r++) row_pixels[r] = tmp.pixel_of_row[r];
    return XPCS_OK;
}

extern "C" int xpcs_create(const XpcsParams *prm, int device, xpcs_handle *out)
{
    if (!out) return fail(nullptr, XPCS_E_ARG, "out is NULL");
    *out = nullptr;
    {
        int rc = check_params(prm);
        if (rc) return rc;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, XPCS_E_CUDA, "no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, XPCS_E_ARG, "device %d out of range (%d devices)", device, ndev);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, XPCS_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10)
        return fail(nullptr, XPCS_E_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device,
                    prop.major, prop.minor);

    xpcs_handle_s *h = new xpcs_handle_s();
    h->prm = *prm;
    h->device = device;
    h->P = prm->width * prm->height;
    h->flat_host.assign(h->P, 1.0);
    h->flat_is_one = true;
    if (prm->flatfield) {
        for (int i = 0; i < h->P; i++) {
            h->flat_host[i] = prm->flatfield[i];
            if (prm->flatfield[i] != 1.0) h->flat_is_one = false;
        }
    }
    int rc = build_schedule(h);
    if (!rc) rc = check_cuda(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "stream");
    h->own_stream = true;
    if (!rc) rc = build_maps(h);
    if (rc) {
        g_create_error = h->err;
        xpcs_destroy(h);
        return rc;
    }
    // the handle keeps no pointer into caller memory
    h->prm.dqmap = nullptr;
    h->prm.sqmap = nullptr;
    h->prm.flatfield = nullptr;
    reset_ingest(h);
    *out = h;
    return XPCS_OK;
}

extern "C" void xpcs_destroy(xpcs_handle h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto &kv : h->stats)
        for (auto &pr : kv.second.pending) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
    comm_destroy(h);
    release(h->d_row_of_pixel); release(h->d_pixel_of_row); release(h->d_sbin_of_row); release(h->d_flat);
    release(h->d_lseg_row_start); release(h->d_seg_dq_all); release(h->d_seg_npix_all);
    release(h->d_dark_avg); release(h->d_dark_std); release(h->d_dense_bound); release(h->d_dense_every); release(h->d_dense_args);
    release(h->d_idx); release(h->d_val); release(h->d_evt); release(h->d_valf); release(h->d_frame_off);
    release(h->d_dense_counter);
    release(h->d_stream_state); release(h->d_st_idx2); release(h->d_st_val2); release(h->d_st_off2);
    for (int b = 0; b < 2; b++)
        if (h->ev_st_copy[b]) cudaEventDestroy(h->ev_st_copy[b]);
    release(h->d_row_count); release(h->d_row_len); release(h->d_slice_len); release(h->d_slice_base);
    release(h->d_slice_cur); release(h->d_slice_rec); release(h->d_slice_end); release(h->d_rec);
    for (int k = 0; k < kMaxChunks; k++) {
        release(h->chunk[k].store); release(h->chunk[k].slice_base); release(h->chunk[k].row_len);
        if (h->ev_chunk[k]) cudaEventDestroy(h->ev_chunk[k]);
    }
    release(h->d_block_first); release(h->d_store); release(h->d_summary); release(h->d_frame_acc);
    release(h->d_row_sum); release(h->d_part_total); release(h->d_part_partial); release(h->d_frame_scale);
    release(h->d_G2); release(h->d_IP); release(h->d_IF); release(h->d_partials); release(h->d_scratch); release(h->d_mt_fallback);
    release(h->d_tt_hi); release(h->d_tt_lo); release(h->d_tt_C); release(h->d_tt_sg); release(h->d_tt_out);
    release(h->d_tt_sgint); release(h->d_tt_diag); release(h->d_tt_tiles);
    if (h->stage) cudaFreeHost(h->stage);
    release(h->d_dense_stage[0]); release(h->d_dense_stage[1]);
    for (int b = 0; b < 2; b++) {
        if (h->ev_copied[b]) cudaEventDestroy(h->ev_copied[b]);
        if (h->ev_filtered[b]) cudaEventDestroy(h->ev_filtered[b]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int xpcs_set_stream(xpcs_handle h, void *s)
{
    if (!h) return XPCS_E_ARG;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    if (s) {
        h->stream = (cudaStream_t)s;
        h->own_stream = false;
    } else {
        h->own_stream = true;
        return check_cuda(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "stream");
    }
    return XPCS_OK;
}

extern "C" int xpcs_reset(xpcs_handle h)
{
    if (!h) return XPCS_E_ARG;
    cudaSetDevice(h->device);
    reset_ingest(h);
    return XPCS_OK;
}

extern "C" int xpcs_get_info(xpcs_handle h, XpcsInfo *info)
{
    if (!h || !info) return XPCS_E_ARG;
    info->n_delays = h->T;
    info->max_level = h->max_level;
    info->n_static = h->S;
    info->n_dynamic = h->Q;
    info->n_segments = h->nseg_total;
    info->n_rows = h->R;
    info->n_rows_total = h->R_total;
    info->raw_frames_seen = h->raw_frames;
    info->events_pushed = h->E;
    info->events_stored = h->events_stored;
    info->store_words = h->store_words;
    info->value_kind = h->kind;
    info->max_row_events = h->max_row;
    return XPCS_OK;
}

extern "C" int xpcs_get_row_pixels(xpcs_handle h, int32_t *out)
{
    if (!h || !out) return XPCS_E_ARG;
    for (int r = 0; r < h->R; r++) out[r] = h->pixel_of_row[r];
    return XPCS_OK;
}

// ---------------------------------------------------------------------------------------
// Filter stage
// ---------------------------------------------------------------------------------------
extern "C" int xpcs_set_dark(xpcs_handle h, const int16_t *frames, int n)
{
    if (!h || !frames || n <= 0) return h ? fail(h, XPCS_E_ARG, "set_dark: bad arguments") : XPCS_E_ARG;
    cudaSetDevice(h->device);
    DevBuf<int16_t> tmp;
    int rc = ensure(h, tmp, (size_t)n * h->P, "dark frames");
    if (rc) return rc;
    rc = check_cuda(h, cudaMemcpyAsync(tmp.p, frames, sizeof(int16_t) * (size_t)n * h->P, cudaMemcpyHostToDevice,
                                       h->stream), "dark H2D");
    if (!rc) rc = launch_dark(h, tmp.p, n);
    if (!rc) rc = check_cuda(h, cudaStreamSynchronize(h->stream), "k_dark");
    release(tmp);
    if (!rc) h->have_dark = true;
    h->dense_bounds_ready = false;
    h->dense_args_ready = false;
    return rc;
}

extern "C" int xpcs_get_dark(xpcs_handle h, double *avg, double *sd)
{
    if (!h) return XPCS_E_ARG;
    if (!h->have_dark) return fail(h, XPCS_E_STATE, "no dark image set");
    cudaSetDevice(h->device);
    if (avg) cudaMemcpyAsync(avg, h->d_dark_avg.p, sizeof(double) * h->P, cudaMemcpyDeviceToHost, h->stream);
    if (sd) cudaMemcpyAsync(sd, h->d_dark_std.p, sizeof(double) * h->P, cudaMemcpyDeviceToHost, h->stream);
    return check_cuda(h, cudaStreamSynchronize(h->stream), "dark D2H");
}

__global__ void k_rebase_offsets(int64_t *off, int n, int64_t delta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off[i] += delta;
}

template <typename T>
static int grow(xpcs_handle_s *h, DevBuf<T> &b, size_t used, size_t need, const char *what)
{
    if (b.n >= need && b.p) return XPCS_OK;
    size_t cap = std::max(need, b.n + b.n / 2);
    T *np = nullptr;
    cudaError_t e = cudaMalloc((void **)&np, std::max<size_t>(cap, 1) * sizeof(T));
    if (e != cudaSuccess) return check_cuda(h, e, what);
    if (b.p && used) cudaMemcpyAsync(np, b.p, used * sizeof(T), cudaMemcpyDeviceToDevice, h->stream);
    if (b.p) {
        cudaStreamSynchronize(h->stream);
        cudaFree(b.p);
    }
    b.p = np;
    b.n = cap;
    return XPCS_OK;
}

static void push_timestamps(xpcs_handle_s *h, const double *clock, const double *ticks, int nframes)
{
    for (int i = 0; i < nframes; i++) {
        h->ts_clock.push_back(clock ? clock[i] : 0.0);
        h->ts_ticks.push_back(ticks ? ticks[i] : 0.0);
    }
}

extern "C" int xpcs_push_sparse(xpcs_handle h, const int32_t *idx, const int16_t *val,
                                const int64_t *frame_offsets, const double *clock, const double *ticks,
                                int nframes)
{
    if (!h || !frame_offsets || nframes < 0) return h ? fail(h, XPCS_E_ARG, "push_sparse: bad arguments") : XPCS_E_ARG;
    if (h->ingest_done) return fail(h, XPCS_E_STATE, "push after finish_ingest (call xpcs_reset first)");
    if (h->stream_on) return fail(h, XPCS_E_STATE, "a stream is open (xpcs_stream_begin): use the xpcs_stream_push_* calls");
    if (h->external_events || (h->dense_source && h->raw_frames > 0))
        return fail(h, XPCS_E_STATE, "cannot mix sparse, dense and device pushes in one ingest");
    cudaSetDevice(h->device);
    const int64_t n = frame_offsets[nframes] - frame_offsets[0];
    if (n < 0 || (n > 0 && (!idx || !val))) return fail(h, XPCS_E_ARG, "push_sparse: bad offsets or NULL payload");
    for (int i = 0; i < nframes; i++)
        if (frame_offsets[i + 1] < frame_offsets[i]) return fail(h, XPCS_E_ARG, "push_sparse: frame offsets not monotone");
    size_t need = (size_t)(h->E + n);
    if (h->E == 0 && h->prm.reserve_events > (int64_t)need) need = (size_t)h->prm.reserve_events;
    int rc;
    if ((rc = grow(h, h->d_idx, (size_t)h->E, need + 8, "event indices"))) return rc;
    if ((rc = grow(h, h->d_val, (size_t)h->E, need + 8, "event values"))) return rc;
    const int raw0 = h->raw_frames;
    const int64_t E0 = h->E;
    // host bookkeeping of a push (done while the copies are already under way when pipelined)
    auto bookkeeping = [&]() {
        h->frame_off_host.reserve(h->frame_off_host.size() + (size_t)nframes);
        for (int i = 0; i < nframes; i++)
            h->frame_off_host.push_back(E0 + (frame_offsets[i + 1] - frame_offsets[0]));
        push_timestamps(h, clock, ticks, nframes);
        h->E += n;
        h->raw_frames += nframes;
    };

    // Pipelined ingest (decided at the first push of an ingest): integer photon counts, no
    // stride/average/frame-sum normalisation, a push large enough to be worth cutting up.
    if (raw0 == 0) {
        int64_t min_events = 4 << 20;
        if (const char *e = getenv("XPCS_PIPELINE_MIN_EVENTS")) min_events = atoll(e);
        h->pipe_on = !getenv("XPCS_NO_PIPELINE") && h->flat_is_one && h->prm.avg_frames == 1 &&
                     h->prm.stride_frames == 1 && !h->prm.normalize_by_framesum &&
                     h->prm.frames <= (1 << (32 - kCountBits)) && n >= min_events;
    }
    if (h->pipe_on && !h->pipe_broken && h->pipe_chunks >= kMaxChunks) h->pipe_broken = true;  // chunk table full
    if (!h->pipe_on || h->pipe_broken) {
        if (n > 0) {
            rc = check_cuda(h, cudaMemcpyAsync(h->d_idx.p + E0, idx + frame_offsets[0], sizeof(int32_t) * n,
                                               cudaMemcpyHostToDevice, h->stream), "idx H2D");
            if (!rc) rc = check_cuda(h, cudaMemcpyAsync(h->d_val.p + E0, val + frame_offsets[0], sizeof(int16_t) * n,
                                                        cudaMemcpyHostToDevice, h->stream), "val H2D");
            if (rc) return rc;
        }
        bookkeeping();
        return XPCS_OK;
    }

    // chunks of about n / K events, cut at frame boundaries
    // (four chunks measured best on C3 and C5: more chunks shorten the exposed last ingest but lengthen the
    // concatenation and pad the chunk stores more)
    int K = (int)std::min<int64_t>(4, std::max<int64_t>(1, n / (8 << 20)));
    if (const char *e = getenv("XPCS_PIPELINE_CHUNKS")) K = std::max(1, atoi(e));
    K = std::min(K, kMaxChunks - h->pipe_chunks);
    if (!h->copy_stream) {
        if ((rc = check_cuda(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking), "copy stream"))) return rc;
    }
    // frame offsets of this push: the caller's array goes first on the copy stream (a small copy queued
    // behind the chunk copies would wait for all of them) and is rebased to absolute event indices on
    // the device
    if ((rc = grow(h, h->d_frame_off, (size_t)h->frame_off_uploaded, (size_t)raw0 + nframes + 1, "frame offsets"))) return rc;
    std::vector<int> cut(1, 0);  // frame cuts relative to this push
    for (int k = 1; k < K; k++) {
        // the last chunk is half the size of the others: its ingest is the only one nothing hides
        const int64_t target = frame_offsets[0] + n * (2 * k) / (2 * K - 1);
        int f = (int)(std::lower_bound(frame_offsets, frame_offsets + nframes + 1, target) - frame_offsets);
        f = std::min(std::max(f, cut.back()), nframes);
        if (f > cut.back() && f < nframes) cut.push_back(f);
    }
    cut.push_back(nframes);
    const int nc = (int)cut.size() - 1;
    const int base_chunk = h->pipe_chunks;
    // every chunk's copy is queued up front on the copy stream (which starts behind whatever the
    // handle's stream has queued: buffer growth, earlier chunks) ...
    cudaEvent_t ev_start = nullptr;
    for (int k = 0; k < nc; k++) {
        cudaEvent_t &ev = h->ev_chunk[base_chunk + k];
        if (!ev && (rc = check_cuda(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "chunk event"))) return rc;
    }
    if ((rc = check_cuda(h, cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming), "event"))) return rc;
    cudaEventRecord(ev_start, h->stream);
    cudaStreamWaitEvent(h->copy_stream, ev_start, 0);
    rc = check_cuda(h, cudaMemcpyAsync(h->d_frame_off.p + raw0, frame_offsets, sizeof(int64_t) * ((size_t)nframes + 1),
                                       cudaMemcpyHostToDevice, h->copy_stream), "frame offsets H2D");
    if (rc) return rc;
    cudaEventRecord(ev_start, h->copy_stream);
    cudaStreamWaitEvent(h->stream, ev_start, 0);
    if (E0 != frame_offsets[0]) {
        LaunchScope ls(h, "k_rebase_offsets");
        k_rebase_offsets<<<(nframes + 256) / 256, 256, 0, h->stream>>>(h->d_frame_off.p + raw0, nframes + 1, E0 - frame_offsets[0]);
    }
    h->frame_off_uploaded = (int64_t)raw0 + nframes + 1;
    for (int k = 0; k < nc && !rc; k++) {
        const int64_t a = frame_offsets[cut[k]] - frame_offsets[0], b = frame_offsets[cut[k + 1]] - frame_offsets[0];
        if (b > a) {
            rc = check_cuda(h, cudaMemcpyAsync(h->d_idx.p + E0 + a, idx + frame_offsets[0] + a, sizeof(int32_t) * (b - a),
                                               cudaMemcpyHostToDevice, h->copy_stream), "idx H2D");
            if (!rc) rc = check_cuda(h, cudaMemcpyAsync(h->d_val.p + E0 + a, val + frame_offsets[0] + a, sizeof(int16_t) * (b - a),
                                                        cudaMemcpyHostToDevice, h->copy_stream), "val H2D");
        }
        cudaEventRecord(h->ev_chunk[base_chunk + k], h->copy_stream);
    }
    bookkeeping();
    // ... and each chunk is ingested on the handle's stream as soon as it has arrived
    for (int k = 0; k < nc && !rc && !h->pipe_broken; k++) {
        cudaStreamWaitEvent(h->stream, h->ev_chunk[base_chunk + k], 0);
        const int r2 = launch_ingest_chunk(h, raw0 + cut[k], raw0 + cut[k + 1]);
        if (r2 == 1) h->pipe_broken = true;  // counts beyond the packed word: classic ingest at finish
        else rc = r2;
    }
    cudaStreamSynchronize(h->copy_stream);  // the caller may reuse its buffers on return
    cudaEventDestroy(ev_start);
    return rc;
}

extern "C" int xpcs_push_sparse_device(xpcs_handle h, const int32_t *d_idx, const int16_t *d_val,
                                       const int64_t *d_frame_offsets, int64_t n_events, int nframes)
{
    if (!h || !d_frame_offsets || nframes < 0 || n_events < 0) return h ? fail(h, XPCS_E_ARG, "push_sparse_device: bad arguments") : XPCS_E_ARG;
    if (h->ingest_done) return fail(h, XPCS_E_STATE, "push after finish_ingest (call xpcs_reset first)");
    if (h->stream_on) return fail(h, XPCS_E_STATE, "a stream is open (xpcs_stream_begin): use the xpcs_stream_push_* calls");
    if (h->raw_frames > 0) return fail(h, XPCS_E_STATE, "push_sparse_device must be the only push of an ingest");
    // the ingest kernels read four events per thread with 16- and 8-byte loads
    if ((reinterpret_cast<uintptr_t>(d_idx) & 15u) || (reinterpret_cast<uintptr_t>(d_val) & 7u) ||
        (reinterpret_cast<uintptr_t>(d_frame_offsets) & 7u))
        return fail(h, XPCS_E_ARG, "push_sparse_device: d_idx must be 16-byte aligned, d_val and d_frame_offsets 8-byte aligned");
    h->external_events = true;
    h->ev_idx = d_idx;
    h->ev_val = d_val;
    h->ev_off = d_frame_offsets;
    h->E = n_events;
    h->raw_frames = nframes;
    h->ts_clock.assign(nframes, 0.0);
    h->ts_ticks.assign(nframes, 0.0);
    return XPCS_OK;
}

// ---- frame-slab pushes (multi-GPU: every GPU gets its share of the FRAMES of the whole detector) ----
static int slab_begin(xpcs_handle_s *h, int first_raw_frame, int nframes, int64_t n_events)
{
    if (h->ingest_done) return fail(h, XPCS_E_STATE, "push after finish_ingest (call xpcs_reset first)");
    if (h->stream_on) return fail(h, XPCS_E_STATE, "a stream is open (xpcs_stream_begin): use the xpcs_stream_push_* calls");
    if (h->raw_frames > 0 || h->slab_mode || h->external_events || h->dense_source)
        return fail(h, XPCS_E_STATE, "a frame slab must be the only push of an ingest");
    if (first_raw_frame < 0 || nframes < 0 || n_events < 0) return fail(h, XPCS_E_ARG, "push_sparse_slab: bad arguments");
    if (h->prm.shard_count > 1 && !comm_active(h))
        return fail(h, XPCS_E_STATE, "push_sparse_slab needs the communicator (xpcs_comm_init) when shard_count > 1");
    if (h->prm.shard_count == 1 && first_raw_frame != 0) return fail(h, XPCS_E_ARG, "a single shard's slab starts at frame 0");
    return XPCS_OK;
}

// one shard: a slab is a plain push (unless a one-rank communicator exists and the tests force the exchange)
static bool slab_is_plain_push(const xpcs_handle_s *h)
{
    return h->prm.shard_count == 1 && !(h->comm && getenv("XPCS_SLAB_FORCE_EXCHANGE"));
}

extern "C" int xpcs_push_sparse_slab_device(xpcs_handle h, int first_raw_frame, const int32_t *d_idx, const int16_t *d_val,
                                            const int64_t *d_frame_offsets, int64_t n_events, int nframes)
{
    if (!h || !d_frame_offsets) return h ? fail(h, XPCS_E_ARG, "push_sparse_slab_device: bad arguments") : XPCS_E_ARG;
    int rc = slab_begin(h, first_raw_frame, nframes, n_events);
    if (rc) return rc;
    if (slab_is_plain_push(h)) return xpcs_push_sparse_device(h, d_idx, d_val, d_frame_offsets, n_events, nframes);
    h->slab_mode = true;
    h->slab_first = first_raw_frame;
    h->slab_frames = nframes;
    h->slab_events = n_events;
    h->slab_idx = d_idx;
    h->slab_val = d_val;
    h->slab_off = d_frame_offsets;
    return XPCS_OK;
}

extern "C" int xpcs_push_sparse_slab(xpcs_handle h, int first_raw_frame, const int32_t *idx, const int16_t *val,
                                     const int64_t *frame_offsets, const double *clock, const double *ticks, int nframes)
{
    if (!h || !frame_offsets) return h ? fail(h, XPCS_E_ARG, "push_sparse_slab: bad arguments") : XPCS_E_ARG;
    const int64_t n = nframes >= 0 ? frame_offsets[nframes] - frame_offsets[0] : -1;
    int rc = slab_begin(h, first_raw_frame, nframes, n);
    if (rc) return rc;
    if (slab_is_plain_push(h)) return xpcs_push_sparse(h, idx, val, frame_offsets, clock, ticks, nframes);
    if (n > 0 && (!idx || !val)) return fail(h, XPCS_E_ARG, "push_sparse_slab: NULL payload");
    for (int i = 0; i < nframes; i++)
        if (frame_offsets[i + 1] < frame_offsets[i]) return fail(h, XPCS_E_ARG, "push_sparse_slab: frame offsets not monotone");
    cudaSetDevice(h->device);
    if ((rc = ensure(h, h->d_slab_idx, (size_t)n + 8, "slab indices"))) return rc;
    if ((rc = ensure(h, h->d_slab_val, (size_t)n + 8, "slab values"))) return rc;
    if ((rc = ensure(h, h->d_slab_off, (size_t)nframes + 1, "slab offsets"))) return rc;
    // (with pinned host memory the three copies are asynchronous; the partition kernels of the exchange queue behind them)
    rc = check_cuda(h, cudaMemcpyAsync(h->d_slab_off.p, frame_offsets, sizeof(int64_t) * ((size_t)nframes + 1),
                                       cudaMemcpyHostToDevice, h->stream), "slab offsets H2D");
    if (!rc && n > 0) {
        rc = check_cuda(h, cudaMemcpyAsync(h->d_slab_idx.p, idx + frame_offsets[0], sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice,
                                           h->stream), "slab idx H2D");
        if (!rc) rc = check_cuda(h, cudaMemcpyAsync(h->d_slab_val.p, val + frame_offsets[0], sizeof(int16_t) * (size_t)n,
                                                    cudaMemcpyHostToDevice, h->stream), "slab val H2D");
    }
    if (rc) return rc;
    if (frame_offsets[0] != 0) {
        LaunchScope ls(h, "k_rebase_offsets");
        k_rebase_offsets<<<(nframes + 256) / 256, 256, 0, h->stream>>>(h->d_slab_off.p, nframes + 1, -frame_offsets[0]);
    }
    h->slab_mode = true;
    h->slab_first = first_raw_frame;
    h->slab_frames = nframes;
    h->slab_events = n;
    h->slab_idx = h->d_slab_idx.p;
    h->slab_val = h->d_slab_val.p;
    h->slab_off = h->d_slab_off.p;
    // timestamps of the slab's frames; the other frames' stay with their ranks (xpcs_get_timestamps)
    h->ts_clock.assign((size_t)first_raw_frame, 0.0);
    h->ts_ticks.assign((size_t)first_raw_frame, 0.0);
    push_timestamps(h, clock, ticks, nframes);
    return XPCS_OK;
}

static int dense_prepare(xpcs_handle_s *h)
{
    if (h->dense_source) return XPCS_OK;
    int rc;
    const int F = h->prm.frames;
    int64_t cap = h->prm.reserve_events > 0 ? h->prm.reserve_events : std::min<int64_t>((int64_t)h->R * F, 1LL << 28);
    cap = (cap + 7) & ~7LL;
    if ((rc = ensure(h, h->d_idx, (size_t)cap, "dense event pixels"))) return rc;
    if ((rc = ensure(h, h->d_evt, (size_t)cap, "dense event frames"))) return rc;
    if ((rc = ensure(h, h->d_valf, (size_t)cap, "dense event values"))) return rc;
    if ((rc = ensure(h, h->d_dense_counter, 1, "dense counter"))) return rc;
    if ((rc = ensure(h, h->d_summary, 8, "summary"))) return rc;
    if ((rc = ensure(h, h->d_frame_acc, (size_t)F, "frame sums"))) return rc;
    cudaMemsetAsync(h->d_dense_counter.p, 0, sizeof(unsigned long long), h->stream);
    cudaMemsetAsync(h->d_summary.p, 0, sizeof(long long) * 8, h->stream);
    cudaMemsetAsync(h->d_frame_acc.p, 0, sizeof(double) * (size_t)F, h->stream);
    h->dense_source = true;
    return XPCS_OK;
}

extern "C" int xpcs_push_dense_device(xpcs_handle h, const int16_t *d_frames, int nframes)
{
    if (!h || !d_frames || nframes <= 0) return h ? fail(h, XPCS_E_ARG, "push_dense_device: bad arguments") : XPCS_E_ARG;
    if (h->ingest_done) return fail(h, XPCS_E_STATE, "push after finish_ingest (call xpcs_reset first)");
    if (h->stream_on) return fail(h, XPCS_E_STATE, "a stream is open (xpcs_stream_begin): use the xpcs_stream_push_* calls");
    if (h->external_events || (!h->dense_source && h->raw_frames > 0))
        return fail(h, XPCS_E_STATE, "cannot mix sparse, dense and device pushes in one ingest");
    cudaSetDevice(h->device);
    int rc = dense_prepare(h);
    if (rc) return rc;
    for (int f0 = 0; f0 < nframes; f0 += 65535 * 64) {  // grid.y limit (64 frames per CTA row)
        const int nb = std::min(65535 * 64, nframes - f0);
        rc = launch_dense_filter(h, d_frames + (size_t)f0 * h->P, h->raw_frames + f0, nb);
        if (rc) return rc;
    }
    h->raw_frames += nframes;
    h->ts_clock.resize(h->raw_frames, 0.0);
    h->ts_ticks.resize(h->raw_frames, 0.0);
    return XPCS_OK;
}

extern "C" int xpcs_push_dense(xpcs_handle h, const int16_t *frames, const double *clock, const double *ticks,
                               int nframes)
{
    if (!h || !frames || nframes <= 0) return h ? fail(h, XPCS_E_ARG, "push_dense: bad arguments") : XPCS_E_ARG;
    if (h->ingest_done) return fail(h, XPCS_E_STATE, "push after finish_ingest (call xpcs_reset first)");
    if (h->stream_on) return fail(h, XPCS_E_STATE, "a stream is open (xpcs_stream_begin): use the xpcs_stream_push_* calls");
    cudaSetDevice(h->device);
    // Two device staging buffers of at most 128 MiB each: batch k+1 crosses PCIe on the copy
    // stream while the filter works on batch k on the handle's stream (events order the reuse
    // of a buffer after the filter that read it).
    const size_t frame_bytes = sizeof(int16_t) * (size_t)h->P;
    const int batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)nframes, (128u << 20) / frame_bytes));
    int rc = XPCS_OK;
    if (!h->copy_stream) rc = check_cuda(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking), "copy stream");
    for (int b = 0; b < 2 && !rc; b++) {
        if (nframes > batch || b == 0) rc = ensure(h, h->d_dense_stage[b], (size_t)batch * h->P, "dense staging");
        if (!rc && !h->ev_copied[b]) rc = check_cuda(h, cudaEventCreateWithFlags(&h->ev_copied[b], cudaEventDisableTiming), "event");
        if (!rc && !h->ev_filtered[b]) rc = check_cuda(h, cudaEventCreateWithFlags(&h->ev_filtered[b], cudaEventDisableTiming), "event");
    }
    if (rc) return rc;
    // the copy stream starts behind whatever the handle's stream has queued so far
    cudaEventRecord(h->ev_filtered[0], h->stream);
    cudaEventRecord(h->ev_filtered[1], h->stream);
    const int before = h->raw_frames;
    int k = 0;
    for (int f0 = 0; f0 < nframes && !rc; f0 += batch, k++) {
        const int nb = std::min(batch, nframes - f0);
        const int b = k & 1;
        cudaStreamWaitEvent(h->copy_stream, h->ev_filtered[b], 0);
        rc = check_cuda(h, cudaMemcpyAsync(h->d_dense_stage[b].p, frames + (size_t)f0 * h->P, frame_bytes * nb,
                                           cudaMemcpyHostToDevice, h->copy_stream), "dense H2D");
        if (rc) break;
        cudaEventRecord(h->ev_copied[b], h->copy_stream);
        cudaStreamWaitEvent(h->stream, h->ev_copied[b], 0);
        rc = xpcs_push_dense_device(h, h->d_dense_stage[b].p, nb);
        cudaEventRecord(h->ev_filtered[b], h->stream);
    }
    cudaStreamSynchronize(h->copy_stream);
    cudaStreamSynchronize(h->stream);  // the caller may reuse `frames` on return
    if (rc) return rc;
    for (int i = 0; i < nframes; i++) {
        h->ts_clock[before + i] = clock ? clock[i] : 0.0;
        h->ts_ticks[before + i] = ticks ? ticks[i] : 0.0;
    }
    return XPCS_OK;
}

// Filter getters (main.cpp:339-343, sparse_filter.cpp:190) in the reference's fp32 operations
__global__ void k_get_frame_sum(const double *__restrict__ facc, float *__restrict__ out, int F, int avg, int P)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    out[f] = (float)(f + 1.0);
    float fs = (float)facc[f];
    if (avg > 1) fs = __fdiv_rn(fs, (float)avg);
    out[F + f] = __fdiv_rn(fs, (float)P);
}

__global__ void k_get_pixel_sum(const double *__restrict__ row_sum, const int *__restrict__ pixel_of_row,
                                float *__restrict__ out, int R, int F)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    out[pixel_of_row[r]] = __fdiv_rn((float)row_sum[r], (float)F);
}

static int filter_getters(xpcs_handle_s *h, float *pixel_sum, float *frame_sum, float *part_total, float *part_partial);

extern "C" int xpcs_finish_ingest(xpcs_handle h, float *pixel_sum, float *frame_sum, float *part_total,
                                  float *part_partial)
{
    if (!h) return XPCS_E_ARG;
    if (h->ingest_done) return fail(h, XPCS_E_STATE, "finish_ingest called twice (call xpcs_reset first)");
    if (h->stream_on) return fail(h, XPCS_E_STATE, "a stream is open: end it with xpcs_stream_finish");
    cudaSetDevice(h->device);
    int rc;
    if (h->prm.normalize_by_framesum && h->prm.shard_count > 1 && !comm_active(h))
        return fail(h, XPCS_E_STATE, "normalize_by_framesum with shard_count > 1 needs the frame sums of all shards: call xpcs_comm_init");
    if (h->slab_mode) {
        std::vector<double> ck, tk;  // the slab's own timestamps survive the exchange
        ck.swap(h->ts_clock);
        tk.swap(h->ts_ticks);
        if ((rc = comm_exchange_slab(h))) return rc;
        for (size_t i = 0; i < ck.size() && i < h->ts_clock.size(); i++) {
            h->ts_clock[i] = ck[i];
            h->ts_ticks[i] = tk[i];
        }
    }
    const bool piped = h->pipe_on && !h->pipe_broken && h->pipe_chunks > 0;
    if (!h->dense_source && !h->external_events && !piped) {
        if ((rc = ensure(h, h->d_frame_off, h->frame_off_host.size(), "frame offsets"))) return rc;
        rc = check_cuda(h, cudaMemcpyAsync(h->d_frame_off.p, h->frame_off_host.data(),
                                           sizeof(int64_t) * h->frame_off_host.size(), cudaMemcpyHostToDevice, h->stream),
                        "frame offsets H2D");
        if (rc) return rc;
        h->ev_idx = h->d_idx.p;
        h->ev_val = h->d_val.p;
        h->ev_off = h->d_frame_off.p;
    }
    if (h->dense_source) {
        if ((rc = ensure(h, h->d_summary, 8, "summary"))) return rc;
    }
    if (piped) rc = launch_ingest_concat(h);
    else rc = launch_ingest(h);
    if (rc) return rc;
    h->ingest_done = true;
    return filter_getters(h, pixel_sum, frame_sum, part_total, part_partial);
}

static int filter_getters(xpcs_handle_s *h, float *pixel_sum, float *frame_sum, float *part_total, float *part_partial)
{
    const int F = h->prm.frames;
    int rc;
    // ---- Filter getters, post-scaled as in main.cpp:339-343 and :360-378 ----
    // pixelSum and frameSum are formed on the device (same fp32 operations) and land in the caller's
    // arrays directly; the small partition sums are finished on the host
    const int S = h->S;
    const int windows_all = (F + h->prm.static_window - 1) / h->prm.static_window;
    const int windows = F / h->prm.static_window;
    if (!pixel_sum && !frame_sum && !part_total && !part_partial) return XPCS_OK;  // nothing to read back
    if (comm_active(h)) {
        // every static bin lives on one shard and frameSum covers all pixels (sparse_filter.cpp:175,190): the
        // element-wise sums over the shards are the single-GPU values (exact for integer counts).  One grouped
        // launch, and only when a caller asks for the sums (every rank must ask for the same ones).
        double *bufs[3];
        size_t counts[3];
        int nb = 0;
        if (frame_sum && !h->frame_acc_reduced) {
            bufs[nb] = h->d_frame_acc.p;
            counts[nb++] = (size_t)F;
            h->frame_acc_reduced = true;
        }
        if ((part_total || part_partial) && !h->part_sums_reduced) {
            bufs[nb] = h->d_part_total.p;
            counts[nb++] = (size_t)S;
            bufs[nb] = h->d_part_partial.p;
            counts[nb++] = (size_t)windows_all * S;
            h->part_sums_reduced = true;
        }
        if (nb > 0 && (rc = comm_allreduce_f64_group(h, bufs, counts, nb))) return rc;
    }
    std::vector<double> ptot(S), ppart((size_t)windows_all * S);
    if (pixel_sum || frame_sum) {
        if ((rc = ensure(h, h->d_scratch, (size_t)h->P + 2 * (size_t)F, "filter getters"))) return rc;
        float *d_ps = h->d_scratch.p, *d_fs = h->d_scratch.p + h->P;
        if (frame_sum) {
            LaunchScope ls(h, "k_get_frame_sum");
            k_get_frame_sum<<<(F + 255) / 256, 256, 0, h->stream>>>(h->d_frame_acc.p, d_fs, F, h->prm.avg_frames, h->P);
        }
        if (pixel_sum) {
            cudaMemsetAsync(d_ps, 0, sizeof(float) * (size_t)h->P, h->stream);
            if (h->R > 0) {
                LaunchScope ls(h, "k_get_pixel_sum");
                k_get_pixel_sum<<<(h->R + 255) / 256, 256, 0, h->stream>>>(h->d_row_sum.p, h->d_pixel_of_row.p, d_ps, h->R, F);
            }
            if ((rc = comm_allreduce_f32(h, d_ps, (size_t)h->P))) return rc;  // every pixel has one owner: zeros elsewhere
            cudaMemcpyAsync(pixel_sum, d_ps, sizeof(float) * (size_t)h->P, cudaMemcpyDeviceToHost, h->stream);
        }
        if (frame_sum) cudaMemcpyAsync(frame_sum, d_fs, sizeof(float) * 2 * (size_t)F, cudaMemcpyDeviceToHost, h->stream);
    }
    if (S > 0) {
        cudaMemcpyAsync(ptot.data(), h->d_part_total.p, sizeof(double) * S, cudaMemcpyDeviceToHost, h->stream);
        cudaMemcpyAsync(ppart.data(), h->d_part_partial.p, sizeof(double) * (size_t)windows_all * S, cudaMemcpyDeviceToHost, h->stream);
    }
    if ((rc = check_cuda(h, cudaStreamSynchronize(h->stream), "filter sums D2H"))) return rc;
    if (part_total)
        for (int i = 0; i < S; i++) {
            const float denom = (float)h->pixels_per_sbin[i] * F;  // main.cpp:375-378
            part_total[i] = (float)ptot[i] / denom;
        }
    if (part_partial)
        for (int j = 0; j < windows; j++)
            for (int i = 0; i < S; i++) {
                const float denom = (float)h->pixels_per_sbin[i] * h->prm.static_window;  // main.cpp:363-373
                part_partial[(size_t)j * S + i] = (float)ppart[(size_t)j * S + i] / denom;
            }
    return XPCS_OK;
}

// --frameout (main.cpp:276-310): one lane per row walks its column while the frame is below nframes
template <int KIND>
__global__ void __launch_bounds__(32) k_get_frames(const void *store, const int64_t *__restrict__ slice_base,
                                                    const int *__restrict__ row_len, const int *__restrict__ pixel_of_row,
                                                    float *__restrict__ out, int nframes, int P, int R)
{
    const int s = blockIdx.x, lane = threadIdx.x;
    const int r = s * kSlice + lane;
    if (r >= R) return;
    const int n = row_len[r];
    const int pix = pixel_of_row[r];
    if (KIND == kPacked) {
        const uint32_t *col = reinterpret_cast<const uint32_t *>(store) + slice_base[s] + lane;
        for (int j = 0; j < n; j++) {
            const uint32_t w = col[(int64_t)j * kSlice];
            const int f = (int)(w >> kCountBits);
            if (f >= nframes) break;
            out[(int64_t)f * P + pix] = (float)(w & ((1u << kCountBits) - 1u));
        }
    } else {
        const unsigned long long *col = reinterpret_cast<const unsigned long long *>(store) + slice_base[s] + lane;
        for (int j = 0; j < n; j++) {
            const unsigned long long w = col[(int64_t)j * kSlice];
            const int f = (int)(w >> 32);
            if (f >= nframes) break;
            out[(int64_t)f * P + pix] = __uint_as_float((uint32_t)w);
        }
    }
}

extern "C" int xpcs_get_frames(xpcs_handle h, int nframes, float *out)
{
    if (!h || !out || nframes <= 0) return h ? fail(h, XPCS_E_ARG, "get_frames: bad arguments") : XPCS_E_ARG;
    if (!h->ingest_done) return fail(h, XPCS_E_STATE, "get_frames before finish_ingest");
    if (h->rows_consumed) return fail(h, XPCS_E_STATE, "the event rows were consumed by multitau; call get_frames before it");
    if (nframes > h->prm.frames) nframes = h->prm.frames;
    cudaSetDevice(h->device);
    const size_t n = (size_t)nframes * h->P;
    int rc = ensure(h, h->d_scratch, n, "frame dump");
    if (rc) return rc;
    cudaMemsetAsync(h->d_scratch.p, 0, sizeof(float) * n, h->stream);
    if (h->n_slices > 0) {
        LaunchScope ls(h, "k_get_frames");
        if (h->kind == kPacked)
            k_get_frames<kPacked><<<h->n_slices, 32, 0, h->stream>>>(h->d_store.p, h->d_slice_base.p, h->d_row_len.p,
                                                                     h->d_pixel_of_row.p, h->d_scratch.p, nframes, h->P, h->R);
        else
            k_get_frames<kFloat><<<h->n_slices, 32, 0, h->stream>>>(h->d_store.p, h->d_slice_base.p, h->d_row_len.p,
                                                                    h->d_pixel_of_row.p, h->d_scratch.p, nframes, h->P, h->R);
    }
    rc = check_cuda(h, cudaMemcpyAsync(out, h->d_scratch.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream), "frames D2H");
    if (!rc) rc = check_cuda(h, cudaStreamSynchronize(h->stream), "frames D2H");
    return rc;
}

extern "C" int xpcs_get_timestamps(xpcs_handle h, double *clock, double *ticks)
{
    if (!h) return XPCS_E_ARG;
    const int n = h->raw_frames;
    for (int i = 0; i < n; i++) {
        if (clock) { clock[i] = i + 1; clock[n + i] = h->ts_clock[i]; }
        if (ticks) { ticks[i] = i + 1; ticks[n + i] = h->ts_ticks[i]; }
    }
    return XPCS_OK;
}

// ---------------------------------------------------------------------------------------
// Online multi-tau: frame streams that do not fit the device (multitau_stream.cu)
// ---------------------------------------------------------------------------------------
extern "C" int xpcs_stream_begin(xpcs_handle h, int chunk_frames)
{
    if (!h) return XPCS_E_ARG;
    if (h->ingest_done || h->raw_frames > 0 || h->slab_mode || h->external_events || h->stream_on)
        return fail(h, XPCS_E_STATE, "stream_begin needs a fresh ingest (call xpcs_reset first)");
    cudaSetDevice(h->device);
    int rc = stream_check(h, chunk_frames);
    if (rc) return rc;
    int k = 0;
    while ((1 << k) < chunk_frames) k++;
    h->stream_k = k;
    if ((rc = launch_stream_begin(h))) return rc;
    h->stream_on = true;
    h->stream_done = false;
    h->stream_chunks = 0;
    h->stream_short_seen = false;
    h->rows_consumed = false;
    return XPCS_OK;
}

// one chunk whose events already sit in d_idx / d_val / d_frame_off (local offsets)
static int stream_consume(xpcs_handle_s *h, int nframes, int64_t nev)
{
    const int K = 1 << h->stream_k;
    const int c = h->stream_chunks;
    int rc = launch_ingest_stream_chunk(h, c * K, nframes, nev, c == 0);
    if (rc == 1) return fail(h, XPCS_E_ARG, "stream mode: a photon count does not fit the packed word (0..4095)");
    if (rc) return rc;
    h->pipe_events = (c == 0 ? 0 : h->pipe_events) + h->events_stored;  // events of unmasked pixels so far
    if ((rc = launch_stream_chunk(h, c, false))) return rc;
    h->stream_chunks = c + 1;
    if (nframes < K) h->stream_short_seen = true;
    h->raw_frames += nframes;
    h->E += nev;
    return XPCS_OK;
}

static int stream_push_checks(xpcs_handle_s *h, int nframes)
{
    if (!h->stream_on) return fail(h, XPCS_E_STATE, "stream_push before xpcs_stream_begin");
    if (h->stream_short_seen) return fail(h, XPCS_E_STATE, "stream_push after a short (final) chunk");
    if (nframes <= 0 || h->raw_frames + (int64_t)nframes > h->prm.frames)
        return fail(h, XPCS_E_ARG, "stream_push: %d frames on top of %d exceed the %d of the job", nframes, h->raw_frames, h->prm.frames);
    return XPCS_OK;
}

extern "C" int xpcs_stream_push_sparse(xpcs_handle h, const int32_t *idx, const int16_t *val, const int64_t *frame_offsets,
                                       const double *clock, const double *ticks, int nframes)
{
    if (!h || !frame_offsets) return h ? fail(h, XPCS_E_ARG, "stream_push_sparse: bad arguments") : XPCS_E_ARG;
    int rc = stream_push_checks(h, nframes);
    if (rc) return rc;
    cudaSetDevice(h->device);
    for (int i = 0; i < nframes; i++)
        if (frame_offsets[i + 1] < frame_offsets[i]) return fail(h, XPCS_E_ARG, "stream_push_sparse: frame offsets not monotone");
    if (frame_offsets[nframes] > frame_offsets[0] && (!idx || !val)) return fail(h, XPCS_E_ARG, "stream_push_sparse: NULL payload");
    const int K = 1 << h->stream_k;
    if (!h->copy_stream && (rc = check_cuda(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking), "copy stream"))) return rc;
    for (int b = 0; b < 2; b++)
        if (!h->ev_st_copy[b] && (rc = check_cuda(h, cudaEventCreateWithFlags(&h->ev_st_copy[b], cudaEventDisableTiming), "event"))) return rc;
    // A push may hold several chunks.  Chunk i + 1 is copied (copy stream) into the spare set of event buffers while
    // chunk i is ingested and folded into the state on the handle's stream; the two sets swap after every chunk.  The
    // spare set is free by then: the ingest that read it two chunks ago ended with a synchronisation of the stream.
    const int nch = (nframes + K - 1) / K;
    auto issue_copy = [&](int i, DevBuf<int32_t> &bi, DevBuf<int16_t> &bv, DevBuf<int64_t> &bo) -> int {
        const int f0 = i * K, nf = std::min(K, nframes - f0);
        const int64_t e0 = frame_offsets[f0], nev = frame_offsets[f0 + nf] - e0;
        int r;
        if ((r = grow(h, bi, 0, (size_t)nev + 8, "event indices"))) return r;
        if ((r = grow(h, bv, 0, (size_t)nev + 8, "event values"))) return r;
        if ((r = grow(h, bo, 0, (size_t)K + 1, "frame offsets"))) return r;
        if (nev > 0) {
            r = check_cuda(h, cudaMemcpyAsync(bi.p, idx + e0, sizeof(int32_t) * nev, cudaMemcpyHostToDevice, h->copy_stream), "idx H2D");
            if (!r) r = check_cuda(h, cudaMemcpyAsync(bv.p, val + e0, sizeof(int16_t) * nev, cudaMemcpyHostToDevice, h->copy_stream), "val H2D");
            if (r) return r;
        }
        r = check_cuda(h, cudaMemcpyAsync(bo.p, frame_offsets + f0, sizeof(int64_t) * ((size_t)nf + 1), cudaMemcpyHostToDevice,
                                          h->copy_stream), "frame offsets H2D");
        if (r) return r;
        return check_cuda(h, cudaEventRecord(h->ev_st_copy[i & 1], h->copy_stream), "copy event");
    };
    rc = issue_copy(0, h->d_idx, h->d_val, h->d_frame_off);
    for (int i = 0; i < nch && !rc; i++) {
        const int f0 = i * K, nf = std::min(K, nframes - f0);
        const int64_t e0 = frame_offsets[f0], nev = frame_offsets[f0 + nf] - e0;
        if (i + 1 < nch && (rc = issue_copy(i + 1, h->d_st_idx2, h->d_st_val2, h->d_st_off2))) break;
        cudaStreamWaitEvent(h->stream, h->ev_st_copy[i & 1], 0);
        if (e0 != 0) {
            LaunchScope ls(h, "k_rebase_offsets");
            k_rebase_offsets<<<(nf + 256) / 256, 256, 0, h->stream>>>(h->d_frame_off.p, nf + 1, -e0);
        }
        push_timestamps(h, clock ? clock + f0 : nullptr, ticks ? ticks + f0 : nullptr, nf);
        rc = stream_consume(h, nf, nev);  // synchronises the handle's stream while it sizes the chunk store
        std::swap(h->d_idx, h->d_st_idx2);
        std::swap(h->d_val, h->d_st_val2);
        std::swap(h->d_frame_off, h->d_st_off2);
    }
    cudaStreamSynchronize(h->copy_stream);  // the caller may reuse its buffers on return
    return rc;
}

extern "C" int xpcs_stream_push_sparse_device(xpcs_handle h, const int32_t *d_idx, const int16_t *d_val,
                                              const int64_t *d_frame_offsets, int64_t n_events, int nframes)
{
    if (!h || !d_frame_offsets || n_events < 0) return h ? fail(h, XPCS_E_ARG, "stream_push_sparse_device: bad arguments") : XPCS_E_ARG;
    int rc = stream_push_checks(h, nframes);
    if (rc) return rc;
    if (nframes > (1 << h->stream_k)) return fail(h, XPCS_E_ARG, "stream_push_sparse_device takes one chunk (<= %d frames) per call", 1 << h->stream_k);
    if (n_events > 0 && (!d_idx || !d_val)) return fail(h, XPCS_E_ARG, "stream_push_sparse_device: NULL payload");
    cudaSetDevice(h->device);
    // the chunk is copied into the handle's own buffers (device to device: 6 bytes per event at HBM rate), which
    // keeps the caller free to refill its buffers at once and the ingest kernels on aligned addresses
    if ((rc = grow(h, h->d_idx, 0, (size_t)n_events + 8, "event indices"))) return rc;
    if ((rc = grow(h, h->d_val, 0, (size_t)n_events + 8, "event values"))) return rc;
    if ((rc = grow(h, h->d_frame_off, 0, (size_t)(1 << h->stream_k) + 1, "frame offsets"))) return rc;
    if (n_events > 0) {
        rc = check_cuda(h, cudaMemcpyAsync(h->d_idx.p, d_idx, sizeof(int32_t) * n_events, cudaMemcpyDeviceToDevice, h->stream), "idx D2D");
        if (!rc) rc = check_cuda(h, cudaMemcpyAsync(h->d_val.p, d_val, sizeof(int16_t) * n_events, cudaMemcpyDeviceToDevice, h->stream), "val D2D");
    }
    if (!rc) rc = check_cuda(h, cudaMemcpyAsync(h->d_frame_off.p, d_frame_offsets, sizeof(int64_t) * ((size_t)nframes + 1),
                                                cudaMemcpyDeviceToDevice, h->stream), "frame offsets D2D");
    if (rc) return rc;
    push_timestamps(h, nullptr, nullptr, nframes);
    return stream_consume(h, nframes, n_events);
}

extern "C" int xpcs_stream_finish(xpcs_handle h, float *pixel_sum, float *frame_sum, float *part_total, float *part_partial)
{
    if (!h) return XPCS_E_ARG;
    if (!h->stream_on) return fail(h, XPCS_E_STATE, "stream_finish before xpcs_stream_begin");
    if (h->raw_frames != h->prm.frames)
        return fail(h, XPCS_E_STATE, "stream_finish: %d of the job's %d frames were pushed", h->raw_frames, h->prm.frames);
    cudaSetDevice(h->device);
    int rc = launch_stream_finish(h);
    if (rc) return rc;
    h->stream_on = false;
    h->stream_done = true;
    h->ingest_done = true;
    h->multitau_done = true;
    h->partials_done = false;
    h->rows_consumed = true;   // there are no rows: xpcs_get_frames / xpcs_twotime need a resident ingest
    h->events_stored = h->pipe_events;
    h->store_words = 0;
    h->max_row = 0;
    if ((rc = check_cuda(h, cudaStreamSynchronize(h->stream), "k_stream_finish"))) return rc;
    return filter_getters(h, pixel_sum, frame_sum, part_total, part_partial);
}

// ---------------------------------------------------------------------------------------
// Correlation
// ---------------------------------------------------------------------------------------
extern "C" int xpcs_multitau(xpcs_handle h, float *G2, float *IP, float *IF)
{
    if (!h) return XPCS_E_ARG;
    if (!h->ingest_done) return fail(h, XPCS_E_STATE, "multitau before finish_ingest");
    if (h->rows_consumed && !h->stream_done) return fail(h, XPCS_E_STATE, "the event rows were consumed by a previous multitau; re-ingest");
    cudaSetDevice(h->device);
    int rc = h->stream_done ? XPCS_OK : launch_multitau(h);  // a finished stream holds its G2 / IP / IF already
    if (rc) return rc;
    h->multitau_done = true;
    h->partials_done = false;
    float *dst[3] = {G2, IP, IF};
    const float *src[3] = {h->d_G2.p, h->d_IP.p, h->d_IF.p};
    for (int k = 0; k < 3; k++) {
        if (!dst[k]) continue;
        const size_t n = (size_t)h->T * h->P;
        if ((rc = ensure(h, h->d_scratch, n, "host-layout staging"))) return rc;
        if ((rc = launch_unpermute(h, src[k], h->d_scratch.p))) return rc;
        rc = check_cuda(h, cudaMemcpyAsync(dst[k], h->d_scratch.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream), "G2/IP/IF D2H");
        if (!rc) rc = check_cuda(h, cudaStreamSynchronize(h->stream), "G2/IP/IF D2H");
        if (rc) return rc;
    }
    return XPCS_OK;
}

// G2 / IP / IF of selected detector pixels, [T][n] each: the tau-major layout of xpcs_multitau restricted to the
// listed columns (pixels this shard does not own, or masked ones, come back as zeros)
__global__ void k_gather_correlators(const float *__restrict__ G2, const float *__restrict__ IP, const float *__restrict__ IF,
                                     const int *__restrict__ row_of_pixel, const int *__restrict__ pixels, int n, int T,
                                     int R_pad, int P, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (i >= n) return;
    const int pix = pixels[i];
    const int r = (unsigned)pix < (unsigned)P ? row_of_pixel[pix] : -1;
    const size_t plane = (size_t)T * n, o = (size_t)t * n + i;
    out[o] = r >= 0 ? G2[(size_t)t * R_pad + r] : 0.0f;
    out[plane + o] = r >= 0 ? IP[(size_t)t * R_pad + r] : 0.0f;
    out[2 * plane + o] = r >= 0 ? IF[(size_t)t * R_pad + r] : 0.0f;
}

extern "C" int xpcs_get_correlators(xpcs_handle h, const int32_t *pixels, int n, float *G2, float *IP, float *IF)
{
    if (!h || !pixels || n <= 0) return h ? fail(h, XPCS_E_ARG, "get_correlators: bad arguments") : XPCS_E_ARG;
    if (!h->multitau_done) return fail(h, XPCS_E_STATE, "get_correlators before multitau");
    cudaSetDevice(h->device);
    const size_t plane = (size_t)h->T * n;
    int rc = ensure(h, h->d_scratch, 3 * plane + (size_t)n + 4, "correlator gather");
    if (rc) return rc;
    int *d_pix = reinterpret_cast<int *>(h->d_scratch.p + 3 * plane);
    rc = check_cuda(h, cudaMemcpyAsync(d_pix, pixels, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream), "pixel list H2D");
    if (rc) return rc;
    {
        LaunchScope ls(h, "k_gather_correlators");
        k_gather_correlators<<<dim3((n + 255) / 256, h->T), 256, 0, h->stream>>>(h->d_G2.p, h->d_IP.p, h->d_IF.p, h->d_row_of_pixel.p,
                                                                                d_pix, n, h->T, h->R_pad, h->P, h->d_scratch.p);
    }
    if (G2) cudaMemcpyAsync(G2, h->d_scratch.p, sizeof(float) * plane, cudaMemcpyDeviceToHost, h->stream);
    if (IP) cudaMemcpyAsync(IP, h->d_scratch.p + plane, sizeof(float) * plane, cudaMemcpyDeviceToHost, h->stream);
    if (IF) cudaMemcpyAsync(IF, h->d_scratch.p + 2 * plane, sizeof(float) * plane, cudaMemcpyDeviceToHost, h->stream);
    return check_cuda(h, cudaStreamSynchronize(h->stream), "correlator gather");
}

extern "C" int xpcs_normalize_partials(xpcs_handle h, void **d_partials, int64_t *count)
{
    if (!h) return XPCS_E_ARG;
    if (!h->multitau_done) return fail(h, XPCS_E_STATE, "normalize before multitau");
    cudaSetDevice(h->device);
    int rc = launch_normalize_partials(h);
    if (rc) return rc;
    // with a communicator the exchange happens here (every segment is owned by one shard, the others hold zeros)
    if ((rc = comm_allreduce_f64(h, h->d_partials.p, (size_t)h->partials_count))) return rc;
    h->partials_done = true;
    if (d_partials) *d_partials = h->d_partials.p;
    if (count) *count = h->partials_count;
    return XPCS_OK;
}

extern "C" int xpcs_normalize_finish(xpcs_handle h, float *g2, float *se)
{
    if (!h) return XPCS_E_ARG;
    if (!h->partials_done) return fail(h, XPCS_E_STATE, "normalize_finish before normalize_partials");
    cudaSetDevice(h->device);
    const size_t n = (size_t)h->T * std::max(h->Q, 1);
    int rc = ensure(h, h->d_scratch, 2 * n, "g2 staging");
    if (rc) return rc;
    if ((rc = launch_normalize_finish(h, h->d_scratch.p, h->d_scratch.p + n))) return rc;
    const size_t bytes = sizeof(float) * (size_t)h->T * h->Q;
    if (g2) cudaMemcpyAsync(g2, h->d_scratch.p, bytes, cudaMemcpyDeviceToHost, h->stream);
    if (se) cudaMemcpyAsync(se, h->d_scratch.p + n, bytes, cudaMemcpyDeviceToHost, h->stream);
    return check_cuda(h, cudaStreamSynchronize(h->stream), "g2 D2H");
}

extern "C" int xpcs_normalize(xpcs_handle h, float *g2, float *se)
{
    int rc = xpcs_normalize_partials(h, nullptr, nullptr);
    if (rc) return rc;
    return xpcs_normalize_finish(h, g2, se);
}

extern "C" int xpcs_twotime_sg(xpcs_handle h, int qbin, int wsize, int method, int average, float *C,
                               float *g2full, float *g2partials, float *sg, int *sg_rows)
{
    if (!h) return XPCS_E_ARG;
    if (!h->ingest_done) return fail(h, XPCS_E_STATE, "twotime before finish_ingest");
    if (h->rows_consumed) return fail(h, XPCS_E_STATE, "the event rows were consumed by multitau; re-ingest");
    cudaSetDevice(h->device);
    return launch_twotime(h, qbin, wsize, method, average, C, g2full, g2partials, sg, sg_rows);
}

extern "C" int xpcs_twotime(xpcs_handle h, int qbin, int wsize, int method, int average, float *C,
                            float *g2full, float *g2partials, float *sg)
{
    return xpcs_twotime_sg(h, qbin, wsize, method, average, C, g2full, g2partials, sg, nullptr);
}

// ---------------------------------------------------------------------------------------
// pinned host memory for callers without a CUDA runtime of their own (the host program, the ctypes mirror):
// copies to and from such buffers are asynchronous and run at full PCIe rate
// ---------------------------------------------------------------------------------------
extern "C" void *xpcs_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void xpcs_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------
// measurement hooks
// ---------------------------------------------------------------------------------------
extern "C" int xpcs_kernel_timing(xpcs_handle h, int enable)
{
    if (!h) return XPCS_E_ARG;
    h->timing = enable != 0;
    return XPCS_OK;
}

extern "C" int64_t xpcs_launch_count(xpcs_handle h) { return h ? h->launches : -1; }

extern "C" int64_t xpcs_multitau_fallback_slices(xpcs_handle h)
{
    if (!h || !h->multitau_done || !h->mt_warp_ran) return -1;
    cudaSetDevice(h->device);
    std::vector<unsigned char> f((size_t)h->n_slices);
    if (h->n_slices == 0) return 0;
    cudaStreamSynchronize(h->stream);
    if (cudaMemcpy(f.data(), h->d_mt_fallback.p, f.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    int64_t n = 0;
    for (unsigned char c : f) n += c ? 1 : 0;
    return n;
}

static void drain_events(xpcs_handle_s *h)
{
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto &kv : h->stats) {
        for (auto &pr : kv.second.pending) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) kv.second.ms += ms;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        kv.second.pending.clear();
    }
}

extern "C" int xpcs_kernel_report(xpcs_handle h, const char **names, double *total_ms, int64_t *launches, int cap)
{
    if (!h) return XPCS_E_ARG;
    drain_events(h);
    int i = 0;
    for (auto &kv : h->stats) {
        if (i < cap) {
            if (names) names[i] = kv.first.c_str();
            if (total_ms) total_ms[i] = kv.second.ms;
            if (launches) launches[i] = kv.second.launches;
        }
        i++;
    }
    return i;
}

extern "C" int xpcs_kernel_report_reset(xpcs_handle h)
{
    if (!h) return XPCS_E_ARG;
    drain_events(h);
    for (auto &kv : h->stats) {
        kv.second.ms = 0.0;
        kv.second.launches = 0;
    }
    h->launches = 0;
    return XPCS_OK;
}
