// multitau_stream_core.h -- the per-row routines of the ONLINE multi-tau (k_stream_chunk, multitau_stream.cu).
//
// frame streams that do not fit the device.  The reference reads every frame before it
// correlates (main.cpp:263-268) and bins the levels offline, in place (corr.cpp:349-390); here the frames arrive in
// chunks of K = 2^k frames, a chunk is correlated as soon as its pixel-major store exists, and only a small per-row
// STATE survives from one chunk to the next.  What is computed is the exact-maths form of multiTau2 (SURVEY.md A.2;
// the compat flag XPCS_COMPAT_STALE_TAIL needs the complete row -- A.4 -- and is refused in stream mode):
//   L_l = F >> l, integer bins c_l(g) = sum of the counts of the frames [g << l, (g + 1) << l), valid iff g < L_l
//   G2num(l, t') = sum over valid g >= t' of c_l(g) * c_l(g - t')          (the LATER bin decides the chunk)
//   total_l = sum over valid g of c_l(g)
//   IPnum(l, t') = total_l - (the last t' valid bins)   IFnum(l, t') = total_l - (the first t' bins)
//   G2 = fdiv(float(G2num) * 4^-l, L_l - t'), IP / IF = fdiv(float(num) * 2^-l, L_l - t')   (corr.cpp:420-424)
// -- the same integers and the same single IEEE division as the resident kernels (multitau_slice_core.h), so a
// streamed job equals the resident one bit for bit whenever the latter runs without the compat flag.
//
// State of a row (32-bit words, one contiguous block per row so that a warp reads and writes it coalesced):
//   g2[T] (u64) | total[NL] (u64) | tail[NL][W] | head[NL][W] | pend[NL]      W = 2 dpl, NL = lastl + 1
//   tail[l] = the last W valid bins of level l seen so far (most recent last; zeros before the first frame),
//   head[l] = the first W bins, pend[l] (l >= k) = counts of the chunks inside the level-l bin that is still open.
// Levels below k see K >> l new bins per chunk: they are formed densely in shared memory (x[]: the tail, then the
// new bins) and halved in place from level to level.  Levels from k on receive one bin every 2^(l-k) chunks.
//
// One WARP works on one row.  The file compiles for the host too (tests/host_mt/mt_stream_host.cpp): there the 32
// lanes of every phase run one after the other, phases in the same order as on the device, where __syncwarp()
// separates them -- tests/test_multitau_stream_core.py checks the result against the oracle on the CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ST_HD __host__ __device__ __forceinline__
#else
#define ST_HD inline
#endif

// Three builds of this file: the device (a warp), the host with the lanes run one after the other
// (mt_stream_host.cpp), and the host with 32 threads playing the lanes of a warp, barriers where the device has
// __syncwarp() and random delays in front of every phase (-DST_EMU_WARP, mt_stream_emu.cpp: a missing synchronisation
// lets a fast lane read what a delayed lane has not written yet, and the result no longer equals the oracle).
#if defined(ST_EMU_WARP)
namespace xpcs {
namespace st {
int emu_lane();
void emu_jitter();
void emu_sync(int line);
unsigned long long emu_sum64(unsigned long long v);
unsigned long long emu_shfl_xor64(unsigned long long v, int off);
}  // namespace st
}  // namespace xpcs
#define ST_WARP 1
#define ST_LANE() (xpcs::st::emu_jitter(), xpcs::st::emu_lane())   // every phase starts after a random delay of its lane
#define ST_SYNC() xpcs::st::emu_sync(__LINE__)   // (the harness can drop the barrier of one source line)
#elif defined(__CUDA_ARCH__)
#define ST_WARP 1
#define ST_LANE() ((int)(threadIdx.x & 31u))
#define ST_SYNC() __syncwarp()
#endif
#if defined(ST_WARP)
#define ST_FOR_LANES for (int lane = ST_LANE(), once_ = 1; once_; once_ = 0)
#define ST_NLANE_SLOTS 1
#else
#define ST_FOR_LANES for (int lane = 0; lane < 32; lane++)
#define ST_SYNC() ((void)0)
#define ST_NLANE_SLOTS 32
#endif

namespace xpcs {
namespace st {

constexpr int kCB = 12;  // count bits of the packed word (== kCountBits)
constexpr uint32_t kCMask = (1u << kCB) - 1u;
constexpr int kS = 32;   // rows per slice = stride of a row's column in the chunk store

// launch-uniform schedule (as multitau_slice_core.h: level 0 has the delays 1..cnt0, level l in 1..lastl the
// level-local delays dpl+1..dpl+count) plus the chunk size
struct StSched {
    int F, T, cnt0, lastl, cnt_last;
    int k;  // log2 of the chunk length in frames
    int ev_num;  // a level takes the walk over the row's events instead of the one over its bins while 2 n <= bins * ev_num and 2 n <= 2^k, the room of the frame list (0: never; 1: the product)
};

template <int DPL>
ST_HD int level_count(const StSched &s, int l)
{
    return l == 0 ? s.cnt0 : (l < s.lastl ? DPL : (l == s.lastl ? s.cnt_last : 0));
}
template <int DPL>
ST_HD int level_first(const StSched &s, int l)
{
    return l == 0 ? 0 : s.cnt0 + (l - 1) * DPL;
}
template <int DPL>
ST_HD int level_lo(int l)
{
    return l == 0 ? 1 : DPL + 1;
}

template <int DPL>
struct Layout {
    static constexpr int W = 2 * DPL;
    static constexpr int XPAD = (W + 3) & ~3;  // words in front of the new bins in x[] (the tail sits right before them)
    static ST_HD int nl(const StSched &s) { return s.lastl + 1; }
    static ST_HD int off_total(const StSched &s) { return 2 * s.T; }
    static ST_HD int off_tail(const StSched &s) { return off_total(s) + 2 * nl(s); }
    static ST_HD int off_head(const StSched &s) { return off_tail(s) + nl(s) * W; }
    static ST_HD int off_pend(const StSched &s) { return off_head(s) + nl(s) * W; }
    static ST_HD int words(const StSched &s) { return (off_pend(s) + nl(s) + 3) & ~3; }  // row stride: 16-byte multiple
    static ST_HD int x_words(const StSched &s) { return XPAD + (1 << s.k); }
    // scratch of a warp: x[] | the chunk-local frames of the row's events, 16 bits each, for the event walk (at most
    // 2^k / 2 events take it) | a copy of the row's state, worked on in place of the global one and written back
    static ST_HD int evf_words(const StSched &s) { return (1 << s.k) / 4; }
    static ST_HD int scratch_words(const StSched &s) { return x_words(s) + evf_words(s) + words(s); }
};

ST_HD float pow2_neg(int e)
{
    union { uint32_t u; float f; } c;
    c.u = (uint32_t)(127 - e) << 23;
    return c.f;
}

ST_HD float scaled_div(float num, int neff)
{
#if defined(__CUDA_ARCH__)
    return neff > 0 ? __fdiv_rn(num, (float)neff) : num;
#else
    return neff > 0 ? num / (float)neff : num;
#endif
}

// valid new bins of level l in chunk c
ST_HD int new_valid(const StSched &s, int c, int l)
{
    const int nb = 1 << (s.k - l);
    const long long g0 = (long long)c * nb;
    long long v = (long long)(s.F >> l) - g0;
    return v < 0 ? 0 : (v > nb ? nb : (int)v);
}

#if defined(__CUDA_ARCH__)
ST_HD unsigned long long warp_sum64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#elif defined(ST_EMU_WARP)
inline unsigned long long warp_sum64(unsigned long long v) { return emu_sum64(v); }
#endif

#if defined(ST_WARP)
ST_HD unsigned long long shfl_xor64(unsigned long long v, int off)
{
#if defined(__CUDA_ARCH__)
    return __shfl_xor_sync(0xffffffffu, v, off);
#else
    return emu_shfl_xor64(v, off);
#endif
}

// Warp sums of CNT (16, 8 or 4) per-lane values at once: at every step a lane keeps the half of its values its own lane
// bit selects and hands the other half to its partner, so 16 values cost 8 + 4 + 2 + 1 + 1 exchanges instead of
// 16 x 5.  Returns the complete sum of value number lane >> SH in every lane (SH = 1 / 2 / 3 for 16 / 8 / 4 values).
template <int CNT>
ST_HD unsigned long long fold_sum(const unsigned long long (&acc)[CNT], int lane)
{
    static_assert(CNT == 16 || CNT == 8 || CNT == 4, "fold_sum: 4, 8 or 16 values");
    unsigned long long a[CNT];
#pragma unroll
    for (int i = 0; i < CNT; i++) a[i] = acc[i];
    int off = 16;
#pragma unroll
    for (int n = CNT / 2; n >= 1; n >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; i++) {
            const unsigned long long keep = up ? a[i + n] : a[i];
            const unsigned long long give = up ? a[i] : a[i + n];
            a[i] = keep + shfl_xor64(give, off);
        }
    }
    unsigned long long v = a[0];
    for (; off >= 1; off >>= 1) v += shfl_xor64(v, off);
    return v;
}
#endif

// ---- a level with many new bins (nb >= 32): every lane takes groups of four consecutive bins (one 16-byte
// shared-memory load per four words, consecutive lanes on consecutive vectors: conflict free) with the W bins
// before them, CNT delays lo..lo+CNT-1 at once.  Bins from nvalid on read as zero as the LATER element of a pair.
template <int DPL, int LO, int CNT>
ST_HD void mac_groups(const uint32_t *xn, int nvalid, int lane, unsigned long long (&acc)[CNT], unsigned long long &tot)
{
    constexpr int NV = (LO + CNT - 1 + 3) / 4;  // vectors in front of the group
    constexpr int WR = 4 * NV;
    for (int q = lane; 4 * q < nvalid; q += 32) {
        const int t0 = 4 * q;
        uint32_t w[WR + 4];
#pragma unroll
        for (int v = 0; v <= NV; v++) {
#if defined(__CUDA_ARCH__)
            const uint4 u = *reinterpret_cast<const uint4 *>(xn + t0 - WR + 4 * v);
            w[4 * v] = u.x; w[4 * v + 1] = u.y; w[4 * v + 2] = u.z; w[4 * v + 3] = u.w;
#else
            for (int i = 0; i < 4; i++) w[4 * v + i] = xn[t0 - WR + 4 * v + i];
#endif
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t cur = (t0 + i < nvalid) ? w[WR + i] : 0u;
            tot += cur;
#pragma unroll
            for (int d = 0; d < CNT; d++) acc[d] += (unsigned long long)cur * w[WR + i - LO - d];
        }
    }
}

// ---- a level with far more bins than the row has events in this chunk: walk the events instead.  An event is the
// head of its bin at level l if the event before it lies in another bin; a head stands for its bin (the dense array
// holds the bin's count), everything else about the level is as in mac_groups.  Cost ~ (2 CNT + 12) per head event
// against (4 CNT + 20) per four bins there.
template <int DPL, int LO, int CNT>
ST_HD void mac_events(const uint32_t *xn, int nvalid, int lane, const uint16_t *evf, int n, int l,
                      unsigned long long (&acc)[CNT], unsigned long long &tot)
{
    for (int j = lane; j < n; j += 32) {
        const int g = (int)evf[j] >> l;
        if (g >= nvalid) continue;  // (the events ascend: the bins from nvalid on hold the last few of the row)
        if (j > 0 && ((int)evf[j - 1] >> l) == g) continue;
        const uint32_t cur = xn[g];
        const uint32_t *xe = xn + g - LO;
        tot += cur;
#pragma unroll
        for (int d = 0; d < CNT; d++) acc[d] += (unsigned long long)cur * xe[-d];
    }
}

// One level below k: tail in, pairs, totals, head / tail out.  xn = x + XPAD holds the nb new bins of the level.
template <int DPL, int LO, int CNT>
ST_HD void level_dense(const StSched &s, int c, int l, uint32_t *x, uint32_t *st, const uint16_t *evf, int n)
{
    typedef Layout<DPL> LY;
    constexpr int W = LY::W;
    uint32_t *xn = x + LY::XPAD;
    const int nb = 1 << (s.k - l);
    const int nvalid = new_valid(s, c, l);
    const int cnt = level_count<DPL>(s, l);
    if (nvalid == 0 || cnt == 0) return;
    unsigned long long *g2 = reinterpret_cast<unsigned long long *>(st) + level_first<DPL>(s, l);
    unsigned long long *total = reinterpret_cast<unsigned long long *>(st + LY::off_total(s)) + l;
    uint32_t *tail = st + LY::off_tail(s) + l * W;
    uint32_t *head = st + LY::off_head(s) + l * W;
    ST_FOR_LANES
    {
        if (lane < W) x[LY::XPAD - W + lane] = tail[lane];
    }
    ST_SYNC();
    if (nb >= 32) {
        ST_FOR_LANES
        {
            unsigned long long acc[CNT], tot = 0;
#pragma unroll
            for (int d = 0; d < CNT; d++) acc[d] = 0;
            if (s.ev_num > 0 && 2ll * n <= (long long)nb * s.ev_num && 2 * n <= (1 << s.k)) mac_events<DPL, LO, CNT>(xn, nvalid, lane, evf, n, l, acc, tot);
            else mac_groups<DPL, LO, CNT>(xn, nvalid, lane, acc, tot);
#if defined(ST_WARP)
            constexpr int SH = CNT == 16 ? 1 : (CNT == 8 ? 2 : 3);  // value d ends up in the lanes d << SH .. (d << SH) + (1 << SH) - 1
            const unsigned long long mine = fold_sum<CNT>(acc, lane);
            tot = warp_sum64(tot);
            if ((lane & ((1 << SH) - 1)) == 0 && (lane >> SH) < cnt) g2[lane >> SH] += mine;
            if (lane == 31) *total += tot;
#else
            for (int d = 0; d < cnt; d++) g2[d] += acc[d];
            *total += tot;
#endif
        }
    } else {
        // few bins: lane d walks all of them for its own delay, lane 31 adds up the total
        ST_FOR_LANES
        {
            if (lane < cnt) {
                unsigned long long acc = 0;
                const uint32_t *xe = xn - (LO + lane);
                for (int t = 0; t < nvalid; t++) acc += (unsigned long long)xn[t] * xe[t];
                g2[lane] += acc;
            } else if (lane == 31) {
                unsigned long long tot = 0;
                for (int t = 0; t < nvalid; t++) tot += xn[t];
                *total += tot;
            }
        }
    }
    // head: the first W bins of the level; tail: the last W of (old tail, new valid bins)
    ST_FOR_LANES
    {
        const long long g0 = (long long)c * nb;
        if (g0 < W)
            for (int t = lane; t < nvalid && g0 + t < W; t += 32) head[g0 + t] = xn[t];
        if (lane < W) tail[lane] = x[LY::XPAD - W + nvalid + lane];
    }
    // (the next writer of x[] -- the halving -- reads, synchronises and only then writes)
}

// x[XPAD + t] = x[XPAD + 2t] + x[XPAD + 2t + 1] for t < nb / 2, in place
template <int DPL>
ST_HD void halve(uint32_t *x, int nb)
{
    uint32_t *xn = x + Layout<DPL>::XPAD;
    const int half = nb >> 1;
    int t0 = 0;
    // 128 outputs per step while they last: a lane adds up eight consecutive words into four (two 16-byte loads, one
    // 16-byte store); a step writes [t0, t0 + 128) and has read [2 t0, 2 t0 + 256): later steps read beyond what it wrote
    for (; t0 + 128 <= half; t0 += 128) {
        uint32_t tmp[ST_NLANE_SLOTS][4];
        ST_FOR_LANES
        {
            const uint32_t *src = xn + 2 * t0 + 8 * lane;
#if defined(__CUDA_ARCH__)
            const uint4 a = *reinterpret_cast<const uint4 *>(src), b = *reinterpret_cast<const uint4 *>(src + 4);
            tmp[0][0] = a.x + a.y; tmp[0][1] = a.z + a.w; tmp[0][2] = b.x + b.y; tmp[0][3] = b.z + b.w;
#else
            for (int i = 0; i < 4; i++) tmp[lane % ST_NLANE_SLOTS][i] = src[2 * i] + src[2 * i + 1];
#endif
        }
        ST_SYNC();
        ST_FOR_LANES
        {
            uint32_t *dst = xn + t0 + 4 * lane;
#if defined(__CUDA_ARCH__)
            *reinterpret_cast<uint4 *>(dst) = make_uint4(tmp[0][0], tmp[0][1], tmp[0][2], tmp[0][3]);
#else
            for (int i = 0; i < 4; i++) dst[i] = tmp[lane % ST_NLANE_SLOTS][i];
#endif
        }
        ST_SYNC();
    }
    for (; t0 < half; t0 += 32) {
        uint32_t tmp[ST_NLANE_SLOTS];
        ST_FOR_LANES
        {
            const int t = t0 + lane;
            tmp[lane % ST_NLANE_SLOTS] = t < half ? xn[2 * t] + xn[2 * t + 1] : 0u;
        }
        ST_SYNC();
        ST_FOR_LANES
        {
            const int t = t0 + lane;
            if (t < half) xn[t] = tmp[lane % ST_NLANE_SLOTS];
        }
        ST_SYNC();
    }
}

// A row without events in this chunk: nothing to add anywhere, the tails of the levels below k move on by the
// level's new valid bins (all zero).
template <int DPL>
ST_HD void tails_skip(const StSched &s, int c, uint32_t *st)
{
    typedef Layout<DPL> LY;
    constexpr int W = LY::W;
    const int lend = s.k < s.lastl + 1 ? s.k : s.lastl + 1;
    for (int l = 0; l < lend; l++) {
        const int nvalid = new_valid(s, c, l);
        if (nvalid == 0) continue;
        uint32_t *tail = st + LY::off_tail(s) + l * W;
        uint32_t tmp[ST_NLANE_SLOTS];
        ST_FOR_LANES
        {
            tmp[lane % ST_NLANE_SLOTS] = (lane < W && lane + nvalid < W) ? tail[lane + nvalid] : 0u;
        }
        ST_SYNC();
        ST_FOR_LANES
        {
            if (lane < W) tail[lane] = tmp[lane % ST_NLANE_SLOTS];
        }
        ST_SYNC();
    }
}

// Levels from k on: the chunk (complete: (c + 1) << k <= F) is one bin of level k with the value S.
template <int DPL>
ST_HD void cascade(const StSched &s, int c, uint32_t S, uint32_t *st)
{
    typedef Layout<DPL> LY;
    constexpr int W = LY::W;
    for (int l = s.k; l <= s.lastl; l++) {
        const int m = l - s.k;
        uint32_t *pend = st + LY::off_pend(s) + l;
        uint32_t *tail = st + LY::off_tail(s) + l * W;
        const bool complete = (((unsigned)(c + 1)) & ((1u << m) - 1u)) == 0u;
        const int g = c >> m;
        const bool valid = g < (s.F >> l);
        const int cnt = level_count<DPL>(s, l), lo = level_lo<DPL>(l);
        uint32_t tmp[ST_NLANE_SLOTS];
        uint32_t vv[ST_NLANE_SLOTS];
        ST_FOR_LANES
        {
            const uint32_t v = *pend + S;
            vv[lane % ST_NLANE_SLOTS] = v;
            if (complete && valid) {
                if (lane < cnt) {
                    unsigned long long *g2 = reinterpret_cast<unsigned long long *>(st) + level_first<DPL>(s, l);
                    g2[lane] += (unsigned long long)v * tail[W - (lo + lane)];
                }
                tmp[lane % ST_NLANE_SLOTS] = lane < W ? (lane + 1 < W ? tail[lane + 1] : v) : 0u;
            }
        }
        ST_SYNC();
        ST_FOR_LANES
        {
            const uint32_t v = vv[lane % ST_NLANE_SLOTS];
            if (complete && valid) {
                if (lane < W) tail[lane] = tmp[lane % ST_NLANE_SLOTS];
                if (lane == 31) {
                    *(reinterpret_cast<unsigned long long *>(st + LY::off_total(s)) + l) += v;
                    if (g < W) st[LY::off_head(s) + l * W + g] = v;
                }
            }
            if (lane == 31) *pend = complete ? 0u : v;
        }
        ST_SYNC();
    }
}

// One chunk of one row.  ev: the row's words of the chunk store (word j at ev[j * 32], frame << 12 | count, absolute
// frame numbers inside [c << k, (c + 1) << k)), n of them; x: Layout::scratch_words() words of scratch, 16-byte aligned;
// gst: the row's state in global memory (16-byte aligned, Layout::words() words).
template <int DPL>
ST_HD void row_chunk(const StSched &s, int c, const uint32_t *ev, int n, uint32_t *x, uint32_t *gst)
{
    typedef Layout<DPL> LY;
    if (n == 0) {  // a few words of the state move: straight in global memory
        tails_skip<DPL>(s, c, gst);
        if (s.k <= s.lastl && ((long long)(c + 1) << s.k) <= s.F) cascade<DPL>(s, c, 0u, gst);
        return;
    }
    const int K = 1 << s.k;
    uint32_t *xn = x + LY::XPAD;
    uint16_t *evf = reinterpret_cast<uint16_t *>(x + LY::x_words(s));
    uint32_t *st = x + LY::x_words(s) + LY::evf_words(s);
    const int nst = LY::words(s);
    const bool keep_frames = s.ev_num > 0 && 2 * n <= K;   // (no level takes the event walk otherwise)
    // the state comes in with full-width loads and is worked on in shared memory: one round trip to global memory per
    // row and chunk instead of two or three per level
    ST_FOR_LANES
    {
#if defined(__CUDA_ARCH__)
        for (int i = lane; i < (LY::XPAD + K) / 4; i += 32) reinterpret_cast<uint4 *>(x)[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = lane; i < nst / 4; i += 32) reinterpret_cast<uint4 *>(st)[i] = reinterpret_cast<const uint4 *>(gst)[i];
#else
        for (int i = lane; i < LY::XPAD + K; i += 32) x[i] = 0u;
        for (int i = lane; i < nst; i += 32) st[i] = gst[i];
#endif
    }
    ST_SYNC();
    ST_FOR_LANES
    {
        const uint32_t fbase = (uint32_t)c << s.k;
        for (int j = lane; j < n; j += 32) {
            const uint32_t w = ev[(int64_t)j * kS];
            const uint32_t f = (w >> kCB) - fbase;
            xn[f] = w & kCMask;
            if (keep_frames) evf[j] = (uint16_t)f;
        }
    }
    ST_SYNC();
    const bool need_sum = s.k <= s.lastl;
    const int lend = need_sum ? s.k : s.lastl + 1;
    for (int l = 0; l < lend; l++) {
        if (l == 0) level_dense<DPL, 1, 2 * DPL>(s, c, 0, x, st, evf, n);
        else level_dense<DPL, DPL + 1, DPL>(s, c, l, x, st, evf, n);
        if (l + 1 < lend || need_sum) halve<DPL>(x, K >> l);
    }
    if (need_sum && ((long long)(c + 1) << s.k) <= s.F) cascade<DPL>(s, c, xn[0], st);
    ST_SYNC();
    ST_FOR_LANES
    {
#if defined(__CUDA_ARCH__)
        for (int i = lane; i < nst / 4; i += 32) reinterpret_cast<uint4 *>(gst)[i] = reinterpret_cast<const uint4 *>(st)[i];
#else
        for (int i = lane; i < nst; i += 32) gst[i] = st[i];
#endif
    }
}

// Result of delay slot ti of a row.
template <int DPL>
ST_HD void row_result(const StSched &s, const uint32_t *st, int ti, float &G2, float &IP, float &IF)
{
    typedef Layout<DPL> LY;
    constexpr int W = LY::W;
    int l, kk;
    if (ti < s.cnt0) {
        l = 0;
        kk = ti;
    } else {
        l = 1 + (ti - s.cnt0) / DPL;
        kk = (ti - s.cnt0) % DPL;
    }
    const int tp = level_lo<DPL>(l) + kk;
    const int neff = (s.F >> l) - tp;
    const unsigned long long g2 = reinterpret_cast<const unsigned long long *>(st)[ti];
    const unsigned long long total = reinterpret_cast<const unsigned long long *>(st + LY::off_total(s))[l];
    const uint32_t *tail = st + LY::off_tail(s) + l * W;
    const uint32_t *head = st + LY::off_head(s) + l * W;
    unsigned long long last = 0, first = 0;
    for (int j = 0; j < tp; j++) {
        last += tail[W - 1 - j];
        first += head[j];
    }
    G2 = scaled_div((float)g2 * pow2_neg(2 * l), neff);
    IP = scaled_div((float)(total - last) * pow2_neg(l), neff);
    IF = scaled_div((float)(total - first) * pow2_neg(l), neff);
}

}  // namespace st
}  // namespace xpcs
