/*
 * xpcs_oracle.c -- CPU restatement of the xpcs-eigen correlation hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (xpcs-eigen_b200/) never links or calls it.
 *
 * Every function restates, in plain C, the arithmetic (operation order and
 * precision included) of one routine of the reference at /root/reference
 * (AdvancedPhotonSource/xpcs-eigen); the reference file:line it follows is cited
 * above each function.  Parity pinning: the reference ships no tests or golden
 * vectors; this restatement is pinned against the reference
 * itself, compiled unmodified into oracle/_ref/corr_ref (see oracle/ref/Makefile) and
 * run on synthetic inputs -- tests/golden/ holds the resulting fixtures and
 * tests/golden/make_golden.py is the script that made them.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * No -ffast-math, no FMA contraction: the reference is a baseline x86-64 build
 * (CMakeLists.txt:5-9), so a*b+c is a rounded multiply followed by a rounded add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* Delay schedule                                                             */
/* ------------------------------------------------------------------------- */

/* corr.cpp:1156-1160 (Corr::calculateLevelMax), double maths as in the reference. */
int xo_level_max(int frames, int dpl)
{
    if (frames < dpl * 2) return 0;
    double a = log2((double)frames);
    double b = log2(1.0 + 1.0 / (double)dpl);
    return (int)(floor(a - b) - log2((double)dpl));
}

/* corr.cpp:1133-1154 (Corr::delaysPerLevel).  Returns the number of delays T and
 * writes at most cap (level, tau) pairs. */
int xo_delay_schedule(int frames, int dpl, int *level, int *tau, int cap)
{
    int max_level = xo_level_max(frames, dpl);
    int n = 0;
    int last = 0;
    for (int lv = 0; lv <= max_level; lv++) {
        int step = (int)pow(2.0, (double)lv);
        int per_level = (lv == 0) ? 2 * dpl : dpl;
        for (int j = 0; j < per_level; j++) {
            if ((double)(last + step) + pow(2.0, (double)lv) > (double)frames) break;
            if (n < cap) {
                level[n] = lv;
                tau[n] = last + step;
            }
            n++;
            last += step;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* Partition maps                                                             */
/* ------------------------------------------------------------------------- */

/*
 * configuration.cpp:244-381 (Configuration::BuildQMap).
 *
 * Output:
 *   mask[P]            1 where dq>0 && sq>0                            (:256-260)
 *   *n_static          max sqmap over valid pixels.  NOTE: the reference only raises
 *                      m_totalStaticPartitions when the dynamic bin already exists
 *                      (:269-274); the two agree whenever the first pixel of a dynamic
 *                      bin is not the only carrier of the largest static id -- the
 *                      precondition SURVEY.md A.8-2 states ("every static bin >= 2 pixels").
 *   *n_dynamic         max dqmap over valid pixels                     (:266-267)
 *   pixels_per_sbin[S] mask pixels per sqmap value                     (:347-356)
 *   Segments (one per surviving (dq, sq) map entry), in std::map iteration order
 *   (dq ascending, sq ascending), pixels ascending inside each (push order):
 *   seg_dq[], seg_sq[], seg_start[nseg+1], seg_pixels[n_mapped].
 *   A static bin that appears under several dynamic bins is kept only under the
 *   dynamic bin holding most of its pixels (first such in ascending dq), with only
 *   that bin's pixels: the reference copies the destination map by value
 *   (:327-328) so the moved pixels are lost.
 * Returns nseg.  Caller sizes seg_* for at most P entries.
 */
int xo_build_qmap(int P, const int *dq, const int *sq, short *mask, int *n_static,
                  int *n_dynamic, int *pixels_per_sbin, int sbin_cap, int *seg_dq,
                  int *seg_sq, int64_t *seg_start, int *seg_pixels)
{
    int S = 0, Q = 0;
    for (int i = 0; i < P; i++) {
        mask[i] = (dq[i] >= 1 && sq[i] >= 1) ? 1 : 0;
        if (mask[i]) {
            if (dq[i] > Q) Q = dq[i];
            if (sq[i] > S) S = sq[i];
        }
    }
    *n_static = S;
    *n_dynamic = Q;
    for (int i = 0; i < sbin_cap; i++) pixels_per_sbin[i] = 0;
    for (int i = 0; i < P; i++)
        if (mask[i] && sq[i] - 1 < sbin_cap) pixels_per_sbin[sq[i] - 1]++;

    /* count[dq][sq] as a dense table indexed (dq-1)*S + (sq-1) */
    int64_t cells = (int64_t)Q * (int64_t)S;
    int64_t *cnt = (int64_t *)calloc((size_t)(cells > 0 ? cells : 1), sizeof(int64_t));
    for (int i = 0; i < P; i++)
        if (mask[i]) cnt[(int64_t)(dq[i] - 1) * S + (sq[i] - 1)]++;

    /* duplicate removal: for each static id keep the dynamic bin with most pixels */
    for (int s = 0; s < S; s++) {
        int owners = 0, best_q = -1;
        int64_t best = 0;
        for (int q = 0; q < Q; q++) {
            int64_t c = cnt[(int64_t)q * S + s];
            if (c > 0) {
                owners++;
                if (c > best) { best = c; best_q = q; }
            }
        }
        if (owners > 1)
            for (int q = 0; q < Q; q++)
                if (q != best_q) cnt[(int64_t)q * S + s] = 0;
    }

    int nseg = 0;
    int64_t off = 0;
    int64_t *cell_start = (int64_t *)malloc((size_t)(cells > 0 ? cells : 1) * sizeof(int64_t));
    for (int q = 0; q < Q; q++)
        for (int s = 0; s < S; s++) {
            int64_t c = cnt[(int64_t)q * S + s];
            cell_start[(int64_t)q * S + s] = -1;
            if (c > 0) {
                seg_dq[nseg] = q + 1;
                seg_sq[nseg] = s + 1;
                seg_start[nseg] = off;
                cell_start[(int64_t)q * S + s] = off;
                off += c;
                nseg++;
            }
        }
    seg_start[nseg] = off;
    for (int i = 0; i < P; i++) {
        if (!mask[i]) continue;
        int64_t cell = (int64_t)(dq[i] - 1) * S + (sq[i] - 1);
        if (cell_start[cell] < 0) continue; /* lost by the duplicate removal */
        seg_pixels[cell_start[cell]++] = i;
    }
    free(cnt);
    free(cell_start);
    return nseg;
}

/* ------------------------------------------------------------------------- */
/* Dark image                                                                 */
/* ------------------------------------------------------------------------- */

/* data_structure/dark_image.cpp:81-106 (DarkImage::Compute): per-pixel running mean
 * and population standard deviation of raw*flatfield over the dark frames, fp64. */
void xo_dark_image(int P, int darks, const int16_t *frames /* [darks][P] */,
                   const double *flat, double *avg, double *std)
{
    for (int j = 0; j < P; j++) { avg[j] = 0.0; std[j] = 0.0; }
    for (int i = 0; i < darks; i++) {
        const int16_t *fr = frames + (int64_t)i * P;
        for (int j = 0; j < P; j++) {
            double before = avg[j];
            /* imm.cpp:98 widens int16 to float first */
            double pix = (double)(float)fr[j] * flat[j];
            avg[j] += (pix - avg[j]) / (double)(i + 1);
            std[j] += (pix - before) * (pix - avg[j]);
        }
    }
    for (int j = 0; j < P; j++) std[j] = sqrt(std[j] / darks);
}

/* ------------------------------------------------------------------------- */
/* Filter stage                                                               */
/* ------------------------------------------------------------------------- */

typedef struct {
    int32_t pix;
    int32_t frame;
    float v;
} xo_event;

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* Shared tail of SparseFilter::Apply (sparse_filter.cpp:164-191) and
 * DenseFilter::Apply (dense_filter.cpp:178-207): emit the touched pixels of one
 * output frame in ascending pixel order and update every accumulator. */
/* io/rigaku.cpp:190-193 moves the static-window number on AFTER the frames 1w, 2w, ... (the
 * Filter stage, sparse_filter.cpp:160-162, before them); the Rigaku restatement switches this on. */
static int g_late_window = 0;
void xo_set_late_window(int on) { g_late_window = on; }

static void emit_frame(int P, int F, int S, int swindow, int frame, int *partition_no,
                       int *touched, int ntouched, float *pix_value, short *touched_map,
                       const int *sbin_of_pixel, float avg_div, float *pixel_sum,
                       float *frame_sum, float *part_total, float *part_partial,
                       xo_event **ev, int64_t *nev, int64_t *cap)
{
    if (!g_late_window && frame > 0 && (frame % swindow) == 0) (*partition_no)++;
    qsort(touched, (size_t)ntouched, sizeof(int), cmp_int);
    float f_sum = 0.0f;
    for (int k = 0; k < ntouched; k++) {
        int pix = touched[k];
        float v = pix_value[pix] / avg_div;
        pixel_sum[pix] += v;
        f_sum += v;
        if (*nev == *cap) {
            *cap = (*cap) * 2 + 1024;
            *ev = (xo_event *)realloc(*ev, (size_t)(*cap) * sizeof(xo_event));
        }
        (*ev)[*nev].pix = pix;
        (*ev)[*nev].frame = frame;
        (*ev)[*nev].v = v;
        (*nev)++;
        int sb = sbin_of_pixel[pix] - 1;
        part_total[sb] += v;
        part_partial[(int64_t)(*partition_no) * S + sb] += v;
        pix_value[pix] = 0.0f;
        touched_map[pix] = 0;
    }
    frame_sum[frame] = (float)(frame + 1.0);
    frame_sum[frame + F] = f_sum / (float)P;
    if (g_late_window && frame > 0 && (frame % swindow) == 0) (*partition_no)++;
}

/* Stable counting sort of the (frame-ascending, pixel-ascending) filtered events
 * into pixel-major rows: the layout of data_structure::SparseData / Row
 * (sparse_data.cpp:59-103, row.h:55-77) flattened to CSR. */
static void events_to_rows(int P, const xo_event *ev, int64_t nev, int64_t *row_ptr,
                           int32_t *row_t, float *row_v)
{
    for (int i = 0; i <= P; i++) row_ptr[i] = 0;
    for (int64_t e = 0; e < nev; e++) row_ptr[ev[e].pix + 1]++;
    for (int i = 0; i < P; i++) row_ptr[i + 1] += row_ptr[i];
    int64_t *cur = (int64_t *)malloc((size_t)P * sizeof(int64_t));
    memcpy(cur, row_ptr, (size_t)P * sizeof(int64_t));
    for (int64_t e = 0; e < nev; e++) {
        int64_t pos = cur[ev[e].pix]++;
        row_t[pos] = ev[e].frame;
        row_v[pos] = ev[e].v;
    }
    free(cur);
}

/*
 * filter/sparse_filter.cpp:115-193 (SparseFilter::Apply) driven by the ingest loop of
 * main.cpp:258-268 and fed by io/imm.cpp:70-118 (Imm::NextFrames).
 *
 * raw frames [0, n_raw) are consumed in blocks of `block` = stride*avg raw frames
 * (main.cpp:258-261); inside a block every stride-th raw frame is used (:143).
 * Outputs (all caller allocated): row_ptr[P+1], row_t/row_v[event_cap] pixel-major
 * rows, pixel_sum[P] (not yet divided by F), frame_sum[2F], part_total[S],
 * part_partial[ceil(F/swindow)*S].  Returns the number of stored events, or -1 when
 * event_cap is too small.
 */
int64_t xo_sparse_filter(int P, int F, int stride, int avg, int swindow, int S,
                         const short *mask, const int *sbin_of_pixel, const double *flat,
                         const int64_t *frame_off, const int32_t *idx, const int16_t *val,
                         int64_t event_cap, int64_t *row_ptr, int32_t *row_t, float *row_v,
                         float *pixel_sum, float *frame_sum, float *part_total,
                         float *part_partial)
{
    int block = stride > 1 ? stride : avg;
    if (stride > 1 && avg > 1) block = stride * avg;
    int windows = (int)ceil((double)F / swindow);
    for (int i = 0; i < P; i++) pixel_sum[i] = 0.0f;
    for (int i = 0; i < S; i++) part_total[i] = 0.0f;
    for (int64_t i = 0; i < (int64_t)windows * S; i++) part_partial[i] = 0.0f;

    float *pix_value = (float *)calloc((size_t)P, sizeof(float));
    short *touched_map = (short *)calloc((size_t)P, sizeof(short));
    int *touched = (int *)malloc((size_t)P * sizeof(int));
    xo_event *ev = NULL;
    int64_t nev = 0, cap = 0;
    int partition_no = 0;

    for (int f = 0; f < F; f++) {
        int ntouched = 0;
        for (int i = 0; i < block; i += stride) {
            int64_t raw = (int64_t)f * block + i;
            for (int64_t e = frame_off[raw]; e < frame_off[raw + 1]; e++) {
                int pix = idx[e];
                if (mask[pix] != 0) {
                    float v = (float)((double)(float)val[e] * flat[pix]);
                    pix_value[pix] += v;
                    if (!touched_map[pix]) { touched_map[pix] = 1; touched[ntouched++] = pix; }
                }
            }
        }
        emit_frame(P, F, S, swindow, f, &partition_no, touched, ntouched, pix_value,
                   touched_map, sbin_of_pixel, (float)avg, pixel_sum, frame_sum, part_total,
                   part_partial, &ev, &nev, &cap);
    }
    int64_t ret = nev;
    if (nev > event_cap) ret = -1;
    else events_to_rows(P, ev, nev, row_ptr, row_t, row_v);
    free(ev); free(pix_value); free(touched_map); free(touched);
    return ret;
}

/*
 * filter/dense_filter.cpp:121-210 (DenseFilter::Apply): per raw frame and unmasked
 * pixel j: v = raw; if darks: v = (float)(v - dark_avg[j]); v = max(v, 0);
 * thresh = (float)(lld + sigma*dark_std[j]); drop when v <= thresh (thresh = 0 without
 * darks); v = (float)(v*flat[j]); accumulate.  dark_avg/dark_std may be NULL.
 * frames points at the first data frame ([n_raw][P] int16).
 */
int64_t xo_dense_filter(int P, int F, int stride, int avg, int swindow, int S,
                        const short *mask, const int *sbin_of_pixel, const double *flat,
                        const double *dark_avg, const double *dark_std, float lld, float sigma,
                        const int16_t *frames, int64_t event_cap, int64_t *row_ptr,
                        int32_t *row_t, float *row_v, float *pixel_sum, float *frame_sum,
                        float *part_total, float *part_partial)
{
    int block = stride > 1 ? stride : avg;
    if (stride > 1 && avg > 1) block = stride * avg;
    int windows = (int)ceil((double)F / swindow);
    for (int i = 0; i < P; i++) pixel_sum[i] = 0.0f;
    for (int i = 0; i < S; i++) part_total[i] = 0.0f;
    for (int64_t i = 0; i < (int64_t)windows * S; i++) part_partial[i] = 0.0f;

    float *pix_value = (float *)calloc((size_t)P, sizeof(float));
    short *touched_map = (short *)calloc((size_t)P, sizeof(short));
    int *touched = (int *)malloc((size_t)P * sizeof(int));
    xo_event *ev = NULL;
    int64_t nev = 0, cap = 0;
    int partition_no = 0;

    for (int f = 0; f < F; f++) {
        int ntouched = 0;
        for (int i = 0; i < block; i += stride) {
            const int16_t *fr = frames + ((int64_t)f * block + i) * P;
            for (int j = 0; j < P; j++) {
                if (mask[j] == 0) continue;
                float v = (float)fr[j];
                float thresh = 0.0f;
                if (dark_avg) {
                    v = (float)((double)v - dark_avg[j]);
                    v = v > 0.0f ? v : 0.0f;
                    thresh = (float)((double)lld + (double)sigma * dark_std[j]);
                }
                if (v <= thresh) continue;
                v = (float)((double)v * flat[j]);
                pix_value[j] += v;
                if (!touched_map[j]) { touched_map[j] = 1; touched[ntouched++] = j; }
            }
        }
        /* dense_filter.cpp:187 divides by the integer average_size_ */
        emit_frame(P, F, S, swindow, f, &partition_no, touched, ntouched, pix_value,
                   touched_map, sbin_of_pixel, (float)avg, pixel_sum, frame_sum, part_total,
                   part_partial, &ev, &nev, &cap);
    }
    int64_t ret = nev;
    if (nev > event_cap) ret = -1;
    else events_to_rows(P, ev, nev, row_ptr, row_t, row_v);
    free(ev); free(pix_value); free(touched_map); free(touched);
    return ret;
}

/*
 * main.cpp:313-343 and :360-378: optional normalise-by-framesum of the stored
 * events, pixel_sum /= F, partition means scaled by pixels_per_sbin*(window|F).
 * Only floor(F/swindow) rows of part_partial are scaled (and later written).
 */
void xo_post_scale(int P, int F, int S, int swindow, int normalize_by_framesum,
                   const int *pixels_per_sbin, const int64_t *row_ptr, const int32_t *row_t,
                   float *row_v, float *pixel_sum, const float *frame_sum, float *part_total,
                   float *part_partial)
{
    if (normalize_by_framesum) {
        float sum = 0.0f;
        for (int i = 0; i < F; i++) sum += frame_sum[i + F];
        float mean = sum / F;
        for (int p = 0; p < P; p++)
            for (int64_t e = row_ptr[p]; e < row_ptr[p + 1]; e++)
                row_v[e] = row_v[e] / (frame_sum[row_t[e] + F] / mean);
    }
    for (int i = 0; i < P; i++) pixel_sum[i] /= F;
    int partitions = (int)floor((double)F / swindow);
    for (int i = 0; i < S; i++)
        for (int j = 0; j < partitions; j++) {
            float denom = (float)pixels_per_sbin[i] * swindow;
            part_partial[(int64_t)j * S + i] /= denom;
        }
    for (int i = 0; i < S; i++) {
        float denom = (float)pixels_per_sbin[i] * F;
        part_total[i] /= denom;
    }
}

/* ------------------------------------------------------------------------- */
/* Multi-tau                                                                  */
/* ------------------------------------------------------------------------- */

/* std::lower_bound as libstdc++ implements it (bits/stl_algobase.h __lower_bound):
 * the probe sequence matters because the reference searches a partly unsorted
 * array. */
static int64_t lower_bound_probe(const int32_t *a, int64_t n, int32_t target)
{
    int64_t first = 0, len = n;
    while (len > 0) {
        int64_t half = len >> 1;
        int64_t mid = first + half;
        if (a[mid] < target) { first = mid + 1; len = len - half - 1; }
        else len = half;
    }
    return first;
}

/*
 * corr.cpp:315-431 (Corr::multiTau2).  Mutates the rows in place exactly as the
 * reference does: level compaction keeps the vectors at their original length, so
 * the binary search of :406 runs over the stale tail too (compat != 0).  With
 * compat == 0 the search is restricted to the live prefix [0, lastIndex), i.e. the
 * mathematically exact sum -- used to quantify the quirk, never for parity.
 * G2/IP/IF are [T][P] (tau-major), pre-zeroed by the caller (main.cpp:193-201).
 * Only pixels that own at least one event are visited (SparseData::ValidPixels).
 */
void xo_multitau(int P, int F, int dpl, const int64_t *row_ptr, int32_t *row_t, float *row_v,
                 float *G2, float *IP, float *IF, int compat, int nthreads)
{
    int cap = 64 * (2 * dpl + 2);
    int *lv = (int *)malloc((size_t)cap * sizeof(int));
    int *tv = (int *)malloc((size_t)cap * sizeof(int));
    int T = xo_delay_schedule(F, dpl, lv, tv, cap);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 20)
    for (int p = 0; p < P; p++) {
        int64_t n0 = row_ptr[p + 1] - row_ptr[p];
        if (n0 == 0) continue;
        int32_t *idx = row_t + row_ptr[p];
        float *val = row_v + row_ptr[p];
        int ll = 0;
        int lastframe = F;
        int64_t live = n0;
        for (int ti = 0; ti < T; ti++) {
            int level = lv[ti];
            int tau = tv[ti];
            if (ll != level) {
                if (lastframe % 2) lastframe -= 1;
                lastframe = lastframe / 2;
                int64_t w = 0, r = 0;
                int i0 = (int)(idx[0] / 2.0), i1;
                if (live > 0 && i0 < lastframe) { idx[0] = i0; r = 1; }
                while (r < live) {
                    i1 = (int)(idx[r] / 2.0);
                    if (i1 >= lastframe) break;
                    if (i1 == i0) val[w] += val[r];
                    else { idx[++w] = i1; val[w] = val[r]; }
                    i0 = i1;
                    r++;
                }
                live = w + 1;
                for (int64_t i = 0; i < live; i++) val[i] /= 2.0f;
            }
            if (level > 0) tau = (int)(tau / pow(2.0, (double)level));
            int64_t o = (int64_t)ti * P + p;
            for (int64_t r = 0; r < live; r++) {
                int src = idx[r];
                if (src < lastframe - tau) {
                    IP[o] += val[r];
                    int64_t pos = lower_bound_probe(idx, compat ? n0 : live, src + tau);
                    if (pos != (compat ? n0 : live) && pos < live && idx[pos] == src + tau)
                        G2[o] += val[r] * val[pos];
                }
                if (src >= tau && src < lastframe) IF[o] += val[r];
            }
            if (lastframe - tau > 0) {
                G2[o] /= (lastframe - tau);
                IP[o] /= (lastframe - tau);
                IF[o] /= (lastframe - tau);
            }
            ll = level;
        }
    }
    free(lv);
    free(tv);
}

/* ------------------------------------------------------------------------- */
/* q-bin normalisation                                                        */
/* ------------------------------------------------------------------------- */

/*
 * corr.cpp:927-1091 (Corr::normalizeG2s) on the partition map of xo_build_qmap.
 * All arithmetic fp32 and sequential in the reference's iteration order: dynamic
 * bin ascending, static bin ascending, pixel ascending.  G2/IP/IF are [T][P].
 * Outputs g2 and stderr in the on-disk layout (T, Q) of norm-0-g2 / norm-0-stderr
 * (h5_result.cpp:80-81 writes the column-major Q x T matrix with dims (T, Q)).
 * Rows of dynamic bins that own no surviving segment stay zero (g2.setZero, :947).
 */
void xo_normalize(int P, int T, int Q, int nseg, const int *seg_dq, const int *seg_sq,
                  const int64_t *seg_start, const int *seg_pixels, const float *G2,
                  const float *IP, const float *IF, float *g2, float *stderr_out)
{
    (void)seg_sq;
    for (int64_t i = 0; i < (int64_t)T * Q; i++) { g2[i] = 0.0f; stderr_out[i] = 0.0f; }
    float *sg = (float *)calloc((size_t)nseg * T, sizeof(float));

    /* :966-1014 static-partition means and their normalised g2 */
    for (int s = 0; s < nseg; s++) {
        int64_t a = seg_start[s], b = seg_start[s + 1];
        int count = (int)(b - a);
        for (int t = 0; t < T; t++) {
            float g = 0.0f, ip = 0.0f, jf = 0.0f;
            for (int64_t k = a; k < b; k++) {
                int64_t o = (int64_t)t * P + seg_pixels[k];
                g += G2[o]; ip += IP[o]; jf += IF[o];
            }
            g /= count; ip /= count; jf /= count;
            sg[(int64_t)s * T + t] = g / (ip * jf);
        }
    }
    /* :1021-1039 NaN-skipping mean over the static bins of each dynamic bin */
    int s0 = 0;
    while (s0 < nseg) {
        int s1 = s0;
        while (s1 < nseg && seg_dq[s1] == seg_dq[s0]) s1++;
        int q = seg_dq[s0];
        for (int t = 0; t < T; t++) {
            float acc = 0.0f, cnt = 0.0f;
            for (int s = s0; s < s1; s++) {
                float x = sg[(int64_t)s * T + t];
                acc += isnan(x) ? 0.0f : x;
                cnt += isnan(x) ? 0.0f : 1.0f;
            }
            g2[(int64_t)t * Q + (q - 1)] = acc / cnt;
        }
        /* :1042-1087 Welford chain over every pixel of the dynamic bin */
        for (int t = 0; t < T; t++) {
            float mean = 0.0f, m2 = 0.0f, n = 0.0f;
            for (int64_t k = seg_start[s0]; k < seg_start[s1]; k++) {
                int64_t o = (int64_t)t * P + seg_pixels[k];
                float before = mean;
                float x = G2[o] / (IP[o] * IF[o]);
                x = isnan(x) ? 0.0f : x;
                n += isnan(x) ? 0.0f : 1.0f;
                float d = x - before;
                float inc = d / n;
                inc = isnan(inc) ? 0.0f : inc;
                mean += inc;
                m2 += (x - before) * (x - mean);
            }
            float norm = m2 / n;
            float inv = 1.0f / n;
            stderr_out[(int64_t)t * Q + (q - 1)] = sqrtf(inv) * sqrtf(norm);
        }
        s0 = s1;
    }
    free(sg);
}

/* ------------------------------------------------------------------------- */
/* Two-time                                                                   */
/* ------------------------------------------------------------------------- */

/*
 * corr.cpp:1166-1226 (ComputeSGSymmetric) + :496-544 (SmoothingSymmetric) and
 * corr.cpp:1228-1305 (ComputeSGStaticMap) + :433-494 (SmoothingStaticMap).
 * bins = the dynamic bins listed in qphi_bin_to_process that exist in the map, in
 * ascending dq order (std::map iteration).  For "symmetric" one sg row per dynamic
 * bin; for "StaticMap" one per static segment of those bins.  Divides the stored
 * values in place.  sg is [nrows][F] or [nrows] when average != 0.  Returns nrows.
 */
int xo_twotime_smooth(int F, int static_map, int average, int nseg, const int *seg_dq,
                      const int64_t *seg_start, const int *seg_pixels, int nproc,
                      const int *qproc, const int64_t *row_ptr, const int32_t *row_t,
                      float *row_v, float *sg)
{
    int nrows = 0;
    int s0 = 0;
    /* pass 1: the sg table */
    int *row_first_seg = (int *)malloc((size_t)(nseg + 1) * sizeof(int));
    int *row_last_seg = (int *)malloc((size_t)(nseg + 1) * sizeof(int));
    while (s0 < nseg) {
        int s1 = s0;
        while (s1 < nseg && seg_dq[s1] == seg_dq[s0]) s1++;
        int hit = 0;
        for (int k = 0; k < nproc; k++) if (qproc[k] == seg_dq[s0]) hit = 1;
        if (hit) {
            if (static_map) {
                for (int s = s0; s < s1; s++) { row_first_seg[nrows] = s; row_last_seg[nrows] = s + 1; nrows++; }
            } else { row_first_seg[nrows] = s0; row_last_seg[nrows] = s1; nrows++; }
        }
        s0 = s1;
    }
    float *full = (float *)calloc((size_t)nrows * F + 1, sizeof(float));
    for (int r = 0; r < nrows; r++) {
        int64_t a = seg_start[row_first_seg[r]], b = seg_start[row_last_seg[r]];
        float *row = full + (int64_t)r * F;
        for (int64_t k = a; k < b; k++) {
            int p = seg_pixels[k];
            for (int64_t e = row_ptr[p]; e < row_ptr[p + 1]; e++) row[row_t[e]] += row_v[e];
        }
        float npix = (float)(b - a);
        for (int f = 0; f < F; f++) row[f] /= npix;
    }
    if (average) {
        for (int r = 0; r < nrows; r++) {
            float acc = 0.0f;
            for (int f = 0; f < F; f++) acc += full[(int64_t)r * F + f];
            sg[r] = acc / (float)F;
        }
    } else memcpy(sg, full, (size_t)nrows * F * sizeof(float));
    /* pass 2: divide the events */
    for (int r = 0; r < nrows; r++) {
        int64_t a = seg_start[row_first_seg[r]], b = seg_start[row_last_seg[r]];
        for (int64_t k = a; k < b; k++) {
            int p = seg_pixels[k];
            for (int64_t e = row_ptr[p]; e < row_ptr[p + 1]; e++) {
                if (average) row_v[e] /= sg[r];
                else row_v[e] /= sg[(int64_t)r * F + row_t[e]];
            }
        }
    }
    free(full); free(row_first_seg); free(row_last_seg);
    return nrows;
}

/*
 * corr.cpp:799-868 (twotimeQBinThreading, per dynamic bin): C[t1][t2] = sum over the
 * bin's pixels of v(t1) v(t2) for t2 >= t1 (row-major, lower triangle zero), divided
 * by the pixel count; g2full[d] = mean of the d-th diagonal; g2partial[d][w] windowed
 * diagonal sums / wsize, d < wsize, w < (F-wsize)/wsize.
 * pix[npix] = the bin's pixel list (all static segments concatenated).
 */
void xo_twotime_bin(int F, int wsize, int npix, const int *pix, const int64_t *row_ptr,
                    const int32_t *row_t, const float *row_v, float *C, float *g2full,
                    float *g2partial)
{
    int partials = (F - wsize) / wsize;
    for (int64_t i = 0; i < (int64_t)F * F; i++) C[i] = 0.0f;
    for (int i = 0; i < F; i++) g2full[i] = 0.0f;
    for (int64_t i = 0; i < (int64_t)wsize * partials; i++) g2partial[i] = 0.0f;
    for (int k = 0; k < npix; k++) {
        int p = pix[k];
        for (int64_t a = row_ptr[p]; a < row_ptr[p + 1]; a++) {
            int64_t f0 = row_t[a];
            float v0 = row_v[a];
            for (int64_t b = a; b < row_ptr[p + 1]; b++)
                C[f0 * F + row_t[b]] += v0 * row_v[b];
        }
    }
    for (int64_t i = 0; i < (int64_t)F * F; i++) C[i] /= npix;
    for (int d = 0; d < F; d++) {
        int count = 0, window = 0;
        for (int x = 0, y = d; x < F - d; x++, y++) {
            g2full[d] += C[(int64_t)x * F + y];
            if (window < partials && d < wsize)
                g2partial[(int64_t)d * partials + window] += C[(int64_t)x * F + y];
            window = (x + 1) / wsize;
            count++;
        }
        g2full[d] /= count;
    }
    for (int64_t i = 0; i < (int64_t)wsize * partials; i++) g2partial[i] /= wsize;
}

int xo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
