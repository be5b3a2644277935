// cabi_recorder.cpp -- TEST SCAFFOLDING, not a product path and not a CPU implementation of anything.
//
// A stand-in for libxpcs_b200.so that COMPUTES NOTHING: it implements the C-ABI symbols the host program `corr`
// imports, records what `corr` hands over (every push: frames, events, the payload arrays, the timestamps; the order
// of the calls) and answers every result request with zeros of the right shape.  tests/test_corr_host_recorder.py
// builds it into a temporary directory next to a copy of the `corr` binary (whose run path is $ORIGIN) and checks,
// on a machine without a GPU, the HOST logic of the drop-in: configuration parsing, the readers, the cut of a frame
// stream into chunks, the call sequence, the names / shapes / types of the result datasets.  It never touches the
// oracle; the numbers `corr` writes in such a run are zeros and nothing compares them with anything.  The product
// library fails loudly without a CUDA device (xpcs_create -> XPCS_E_CUDA); this file is the only other implementer of
// these symbols and lives under tests/.
//
// Record (directory $XPCS_RECORD_DIR, one sub-directory shard<r> per handle of a sharded job; written by xpcs_destroy):
// calls.txt (one line per call), idx.bin / val.bin (sparse payloads in push order), frame_events.bin (int64 events per
// raw frame), clock.bin / ticks.bin (doubles per raw frame), dark.bin / dense.bin (the raw int16 frames of
// xpcs_set_dark / xpcs_push_dense).  Two-time calls leave the caller's arrays untouched.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <string>
#include <vector>

#include "../../include/xpcs_b200.h"

struct xpcs_handle_s {
    XpcsParams prm;
    int P = 0, S = 0, Q = 0, T = 0, rows = 0;
    std::vector<std::string> calls;
    std::vector<int32_t> idx;
    std::vector<int16_t> val;
    std::vector<int64_t> frame_events;
    std::vector<double> clock, ticks;
    bool stream = false;
    int stream_k = 0;
    std::vector<int16_t> dark, dense;      // raw frames handed to set_dark / push_dense
    int slab_first = -1;                   // push_sparse_slab: first raw frame of this handle's slab
    std::string err;
};

static std::string g_err;

static int level_max(int F, int dpl)
{
    if (F < 2 * dpl) return 0;
    return (int)(floor(log2((double)F) - log2(1.0 + 1.0 / (double)dpl)) - log2((double)dpl));
}

static int schedule(int F, int dpl, int32_t *level, int32_t *tau, int cap)
{
    const int ml = level_max(F, dpl);
    int n = 0;
    long long ll = 0;
    for (int i = 0; i <= ml; i++) {
        const long long step = 1ll << i;
        const int ni = i == 0 ? 2 * dpl : dpl;
        for (int j = 0; j < ni; j++) {
            if (ll + step + step > F) break;
            if (n < cap) {
                if (level) level[n] = i;
                if (tau) tau[n] = (int32_t)(ll + step);
            }
            n++;
            ll += step;
        }
    }
    return n;
}

static void say(xpcs_handle_s *h, const char *fmt, long long a = 0, long long b = 0, long long c = 0)
{
    char buf[256];
    snprintf(buf, sizeof(buf), fmt, a, b, c);
    h->calls.push_back(buf);
}

static int record_push(xpcs_handle_s *h, const int32_t *idx, const int16_t *val, const int64_t *off, const double *clock,
                       const double *ticks, int nframes)
{
    for (int f = 0; f < nframes; f++) {
        if (off[f + 1] < off[f]) return XPCS_E_ARG;
        h->frame_events.push_back(off[f + 1] - off[f]);
        h->clock.push_back(clock ? clock[f] : 0.0);
        h->ticks.push_back(ticks ? ticks[f] : 0.0);
    }
    h->idx.insert(h->idx.end(), idx + off[0], idx + off[nframes]);
    h->val.insert(h->val.end(), val + off[0], val + off[nframes]);
    return XPCS_OK;
}

extern "C" {

int xpcs_level_max(int frames, int dpl) { return level_max(frames, dpl); }
int xpcs_delay_schedule(int frames, int dpl, int32_t *level, int32_t *tau, int cap) { return schedule(frames, dpl, level, tau, cap); }

int xpcs_create(const XpcsParams *p, int device, xpcs_handle *out)
{
    if (!p || p->struct_size != (int32_t)sizeof(XpcsParams) || !out) {
        g_err = "recorder: bad XpcsParams";
        return XPCS_E_ARG;
    }
    xpcs_handle_s *h = new xpcs_handle_s;
    h->prm = *p;
    h->P = p->width * p->height;
    for (int i = 0; i < h->P; i++) {
        if (p->dqmap[i] > h->Q) h->Q = p->dqmap[i];
        if (p->sqmap[i] > h->S) h->S = p->sqmap[i];
        if (p->dqmap[i] > 0 && p->sqmap[i] > 0) h->rows++;
    }
    h->T = schedule(p->frames, p->delays_per_level, nullptr, nullptr, 0);
    say(h, "create device=%lld frames=%lld compat_flags=%lld", device, p->frames, p->compat_flags);
    *out = h;
    return XPCS_OK;
}

void xpcs_destroy(xpcs_handle h)
{
    if (!h) return;
    if (const char *dir = getenv("XPCS_RECORD_DIR")) {
        std::string d(dir);
        if (h->prm.shard_count > 1) {   // one handle per shard (corr --gpus N): a record each
            d += "/shard" + std::to_string(h->prm.shard_index);
            mkdir(d.c_str(), 0777);
        }
        auto dump = [&](const char *name, const void *p, size_t bytes) {
            FILE *f = fopen((d + "/" + name).c_str(), "wb");
            if (f) {
                if (bytes) fwrite(p, 1, bytes, f);
                fclose(f);
            }
        };
        std::string text;
        for (const std::string &c : h->calls) text += c + "\n";
        dump("calls.txt", text.data(), text.size());
        dump("idx.bin", h->idx.data(), h->idx.size() * sizeof(int32_t));
        dump("val.bin", h->val.data(), h->val.size() * sizeof(int16_t));
        dump("frame_events.bin", h->frame_events.data(), h->frame_events.size() * sizeof(int64_t));
        dump("clock.bin", h->clock.data(), h->clock.size() * sizeof(double));
        dump("ticks.bin", h->ticks.data(), h->ticks.size() * sizeof(double));
        dump("dark.bin", h->dark.data(), h->dark.size() * sizeof(int16_t));
        dump("dense.bin", h->dense.data(), h->dense.size() * sizeof(int16_t));
    }
    delete h;
}

const char *xpcs_last_error(xpcs_handle h) { return h ? h->err.c_str() : g_err.c_str(); }

int xpcs_get_info(xpcs_handle h, XpcsInfo *info)
{
    if (!h || !info) return XPCS_E_ARG;
    memset(info, 0, sizeof(*info));
    info->n_delays = h->T;
    info->max_level = level_max(h->prm.frames, h->prm.delays_per_level);
    info->n_static = h->S;
    info->n_dynamic = h->Q;
    info->n_rows = info->n_rows_total = h->rows;
    info->raw_frames_seen = (int32_t)h->frame_events.size();
    info->events_pushed = (int64_t)h->idx.size();
    return XPCS_OK;
}

int xpcs_push_sparse(xpcs_handle h, const int32_t *idx, const int16_t *val, const int64_t *off, const double *clock,
                     const double *ticks, int nframes)
{
    if (!h || h->stream) return XPCS_E_STATE;
    say(h, "push_sparse nframes=%lld events=%lld", nframes, off[nframes] - off[0]);
    return record_push(h, idx, val, off, clock, ticks, nframes);
}

int xpcs_finish_ingest(xpcs_handle h, float *pixel_sum, float *frame_sum, float *part_total, float *part_partial)
{
    if (!h || h->stream) return XPCS_E_STATE;
    say(h, "finish_ingest");
    const int F = h->prm.frames, W = F / h->prm.static_window;
    if (pixel_sum) memset(pixel_sum, 0, sizeof(float) * (size_t)h->P);
    if (frame_sum) memset(frame_sum, 0, sizeof(float) * 2 * (size_t)F);
    if (part_total) memset(part_total, 0, sizeof(float) * (size_t)h->S);
    if (part_partial) memset(part_partial, 0, sizeof(float) * (size_t)W * h->S);
    return XPCS_OK;
}

int xpcs_stream_begin(xpcs_handle h, int chunk_frames)
{
    if (!h || h->stream || !h->frame_events.empty()) return XPCS_E_STATE;
    if (h->prm.compat_flags & XPCS_COMPAT_STALE_TAIL) {
        h->err = "recorder: stream mode refuses XPCS_COMPAT_STALE_TAIL, as the library does";
        return XPCS_E_ARG;
    }
    if (chunk_frames < 64 || chunk_frames > 8192 || (chunk_frames & (chunk_frames - 1))) return XPCS_E_ARG;
    h->stream = true;
    h->stream_k = chunk_frames;
    say(h, "stream_begin chunk_frames=%lld", chunk_frames);
    return XPCS_OK;
}

int xpcs_stream_push_sparse(xpcs_handle h, const int32_t *idx, const int16_t *val, const int64_t *off, const double *clock,
                            const double *ticks, int nframes)
{
    if (!h || !h->stream) return XPCS_E_STATE;
    // the library's rule: a push starts on a chunk boundary, only the last push of the job may end inside a chunk
    if (h->frame_events.size() % (size_t)h->stream_k) {
        h->err = "recorder: push after a short chunk";
        return XPCS_E_STATE;
    }
    if ((long long)h->frame_events.size() + nframes > h->prm.frames) return XPCS_E_ARG;
    say(h, "stream_push_sparse nframes=%lld events=%lld first_offset=%lld", nframes, off[nframes] - off[0], off[0]);
    return record_push(h, idx, val, off, clock, ticks, nframes);
}

int xpcs_stream_finish(xpcs_handle h, float *pixel_sum, float *frame_sum, float *part_total, float *part_partial)
{
    if (!h || !h->stream) return XPCS_E_STATE;
    if ((int)h->frame_events.size() != h->prm.frames) {
        h->err = "recorder: stream_finish before all frames were pushed";
        return XPCS_E_STATE;
    }
    h->stream = false;
    int rc = xpcs_finish_ingest(h, pixel_sum, frame_sum, part_total, part_partial);
    h->calls.back() = "stream_finish";
    return rc;
}

int xpcs_get_timestamps(xpcs_handle h, double *clock, double *ticks)
{
    if (!h) return XPCS_E_ARG;
    const size_t n = h->frame_events.size();
    for (size_t i = 0; i < n; i++) {
        if (clock) { clock[i] = (double)(i + 1); clock[n + i] = h->clock[i]; }
        if (ticks) { ticks[i] = (double)(i + 1); ticks[n + i] = h->ticks[i]; }
    }
    return XPCS_OK;
}

int xpcs_multitau(xpcs_handle h, float *G2, float *IP, float *IF)
{
    if (!h) return XPCS_E_ARG;
    say(h, "multitau g2out=%lld", G2 ? 1 : 0);
    const size_t n = (size_t)h->T * h->P;
    if (G2) memset(G2, 0, sizeof(float) * n);
    if (IP) memset(IP, 0, sizeof(float) * n);
    if (IF) memset(IF, 0, sizeof(float) * n);
    return XPCS_OK;
}

int xpcs_normalize(xpcs_handle h, float *g2, float *se)
{
    if (!h) return XPCS_E_ARG;
    say(h, "normalize");
    const size_t n = (size_t)h->T * (h->Q > 0 ? h->Q : 1);
    if (g2) memset(g2, 0, sizeof(float) * n);
    if (se) memset(se, 0, sizeof(float) * n);
    return XPCS_OK;
}

// (the product hands out page-locked memory here)
void *xpcs_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void xpcs_host_free(void *p) { free(p); }

int xpcs_set_dark(xpcs_handle h, const int16_t *frames, int n)
{
    if (!h || !frames || n <= 0) return XPCS_E_ARG;
    say(h, "set_dark n=%lld", n);
    h->dark.assign(frames, frames + (size_t)n * h->P);
    return XPCS_OK;
}

int xpcs_get_dark(xpcs_handle h, double *avg, double *sd)
{
    if (!h || h->dark.empty()) return XPCS_E_STATE;
    say(h, "get_dark");
    if (avg) memset(avg, 0, sizeof(double) * (size_t)h->P);
    if (sd) memset(sd, 0, sizeof(double) * (size_t)h->P);
    return XPCS_OK;
}

int xpcs_push_dense(xpcs_handle h, const int16_t *frames, const double *clock, const double *ticks, int nframes)
{
    if (!h || h->stream || !frames) return XPCS_E_STATE;
    say(h, "push_dense nframes=%lld", nframes);
    h->dense.insert(h->dense.end(), frames, frames + (size_t)nframes * h->P);
    for (int f = 0; f < nframes; f++) {
        h->frame_events.push_back(h->P);
        h->clock.push_back(clock ? clock[f] : 0.0);
        h->ticks.push_back(ticks ? ticks[f] : 0.0);
    }
    return XPCS_OK;
}

int xpcs_get_frames(xpcs_handle h, int nframes, float *out)
{
    if (!h || !out || nframes <= 0) return XPCS_E_ARG;
    say(h, "get_frames n=%lld", nframes);
    memset(out, 0, sizeof(float) * (size_t)nframes * h->P);
    return XPCS_OK;
}

int xpcs_plan_shard(const XpcsParams *p, XpcsShardPlan *plan, int32_t *, int64_t)
{
    if (!p || !plan) return XPCS_E_ARG;
    memset(plan, 0, sizeof(*plan));
    const int P = p->width * p->height;
    for (int i = 0; i < P; i++) {
        if (p->dqmap[i] > plan->n_dynamic) plan->n_dynamic = p->dqmap[i];
        if (p->sqmap[i] > plan->n_static) plan->n_static = p->sqmap[i];
        if (p->dqmap[i] > 0 && p->sqmap[i] > 0) plan->n_rows_total++;
    }
    plan->n_delays = schedule(p->frames, p->delays_per_level, nullptr, nullptr, 0);
    return XPCS_OK;
}

int xpcs_comm_unique_id(void *id128)
{
    memset(id128, 0x5a, 128);
    return XPCS_OK;
}

int xpcs_comm_init(xpcs_handle h, int nranks, int rank, const void *id128)
{
    if (!h || !id128 || nranks != h->prm.shard_count || rank != h->prm.shard_index) return XPCS_E_ARG;
    say(h, "comm_init nranks=%lld rank=%lld id0=%lld", nranks, rank, ((const unsigned char *)id128)[0]);
    return XPCS_OK;
}

int xpcs_push_sparse_slab(xpcs_handle h, int first_raw_frame, const int32_t *idx, const int16_t *val, const int64_t *off,
                          const double *clock, const double *ticks, int nframes)
{
    if (!h || h->stream || h->slab_first >= 0) return XPCS_E_STATE;
    h->slab_first = first_raw_frame;
    say(h, "push_sparse_slab first=%lld nframes=%lld events=%lld", first_raw_frame, nframes, off[nframes] - off[0]);
    return record_push(h, idx, val, off, clock, ticks, nframes);
}

int xpcs_twotime_sg(xpcs_handle h, int qbin, int wsize, int method, int average, float *, float *, float *, float *, int *sg_rows)
{
    if (!h || qbin < 1 || qbin > h->Q) return XPCS_E_ARG;
    say(h, "twotime qbin=%lld wsize=%lld method_average=%lld", qbin, wsize, method * 10 + average);
    if (sg_rows) {
        *sg_rows = 1;
        if (method == 2) {  // StaticMap: one sg row per static partition of the dynamic bin
            std::vector<char> seen((size_t)h->S + 1, 0);
            int n = 0;
            for (int i = 0; i < h->P; i++)
                if (h->prm.dqmap[i] == qbin && h->prm.sqmap[i] > 0 && !seen[(size_t)h->prm.sqmap[i]]) {
                    seen[(size_t)h->prm.sqmap[i]] = 1;
                    n++;
                }
            *sg_rows = n;
        }
    }
    return XPCS_OK;   // the caller's arrays stay as it made them
}

}  // extern "C"
