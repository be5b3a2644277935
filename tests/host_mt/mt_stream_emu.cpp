// mt_stream_emu.cpp -- test infrastructure: the WARP build of multitau_stream_core.h on the CPU.  32 host threads
// play the lanes of one warp: every thread runs row_chunk() exactly as a lane of k_stream_chunk does, a pthread
// barrier stands where the device has __syncwarp(), the 64-bit warp sum goes through a shared table between two
// barriers, and every phase of every lane starts after a random delay.  A phase that reads what another lane writes
// without a barrier in between then sees stale data sooner or later and the result differs from the oracle's
// (tests/test_multitau_stream_core.py checks that removing a barrier is indeed caught).  ThreadSanitizer is used on
// top where it is available; it does not see every such race with 32 threads and 4 shadow slots per word.  The scratch x[] is shared-memory-like: one array for the warp.
#include <assert.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#define ST_EMU_WARP 1
#include "../../xpcs-eigen_b200/csrc/multitau_stream_core.h"

namespace xpcs {
namespace st {
static thread_local int tl_lane = -1;
static pthread_barrier_t g_bar;
static unsigned long long g_tab[32];
static thread_local uint64_t tl_rng = 0;
int emu_lane() { return tl_lane; }
// a lane is held back for a random time before every phase (xorshift per thread): mostly nothing, sometimes a few
// microseconds, now and then a yield -- enough for the other lanes to run a whole phase ahead if nothing stops them
void emu_jitter()
{
    uint64_t x = tl_rng;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    tl_rng = x;
    const unsigned r = (unsigned)(x >> 33) & 15u;
    if (r < 10) return;
    if (r == 15) { std::this_thread::yield(); return; }
    volatile unsigned spin = (r - 9) * 400u;
    while (spin) spin = spin - 1;
}
static int g_drop_line = 0;          // the barrier of this source line of the core is skipped by every lane (0: none)
static int g_sites[64], g_nsites = 0;
void emu_sync(int line)
{
    if (tl_lane == 0 && line > 0) {
        bool seen = false;
        for (int i = 0; i < g_nsites; i++) seen = seen || g_sites[i] == line;
        if (!seen && g_nsites < 64) g_sites[g_nsites++] = line;
    }
    if (line > 0 && line == g_drop_line) return;
    pthread_barrier_wait(&g_bar);
}
unsigned long long emu_sum64(unsigned long long v)
{
    g_tab[tl_lane] = v;
    pthread_barrier_wait(&g_bar);
    unsigned long long s = 0;
    for (int i = 0; i < 32; i++) s += g_tab[i];
    pthread_barrier_wait(&g_bar);
    return s;
}
unsigned long long emu_shfl_xor64(unsigned long long v, int off)
{
    g_tab[tl_lane] = v;
    pthread_barrier_wait(&g_bar);
    const unsigned long long r = g_tab[tl_lane ^ off];
    pthread_barrier_wait(&g_bar);
    return r;
}
}  // namespace st
}  // namespace xpcs

using namespace xpcs::st;

template <int DPL>
static int run(const StSched &sc, int nrows, const int64_t *row_ptr, const int32_t *frames, const int32_t *counts,
               float *G2, float *IP, float *IF)
{
    typedef Layout<DPL> LY;
    static int seed = 0;
    seed += 7919;
    const int K = 1 << sc.k;
    const int nchunks = (sc.F + K - 1) / K;
    const int stride = LY::words(sc);
    std::vector<uint32_t> state((size_t)nrows * stride, 0u);
    std::vector<uint32_t> xs((size_t)LY::scratch_words(sc) + 4, 0xdeadbeefu);
    uint32_t *xa = xs.data();
    while (reinterpret_cast<uintptr_t>(xa) & 15u) xa++;
    // the chunk stores, built up front: word j of row r of chunk c at ev[c][r][j * 32]
    std::vector<std::vector<std::vector<uint32_t>>> ev(nchunks, std::vector<std::vector<uint32_t>>(nrows));
    for (int r = 0; r < nrows; r++)
        for (int64_t j = row_ptr[r]; j < row_ptr[r + 1]; j++) {
            if (counts[j] < 0 || counts[j] > 4095) return 2;
            std::vector<uint32_t> &e = ev[frames[j] >> sc.k][r];
            e.resize(e.size() + 32, 0xffffffffu);
            e[e.size() - 32] = ((uint32_t)frames[j] << kCB) | (uint32_t)counts[j];
        }
    pthread_barrier_init(&g_bar, nullptr, 32);
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; lane++)
        lanes.emplace_back([&, lane]() {
            tl_lane = lane;
            tl_rng = 0x9e3779b97f4a7c15ull * (uint64_t)(lane + 1) + (uint64_t)seed;
            for (int c = 0; c < nchunks; c++)
                for (int r = 0; r < nrows; r++) {
                    row_chunk<DPL>(sc, c, ev[c][r].data(), (int)(ev[c][r].size() / 32), xa, state.data() + (size_t)r * stride);
                    emu_sync(-1);  // the next row reuses x[] (on the device: another warp's own scratch, or the next launch)
                }
        });
    for (auto &t : lanes) t.join();
    pthread_barrier_destroy(&g_bar);
    for (int r = 0; r < nrows; r++)
        for (int ti = 0; ti < sc.T; ti++) {
            const size_t o = (size_t)ti * nrows + r;
            row_result<DPL>(sc, state.data() + (size_t)r * stride, ti, G2[o], IP[o], IF[o]);
        }
    return 0;
}

extern "C" int mt_stream_emu(int dpl, int F, int T, int cnt0, int lastl, int cnt_last, int k, int ev_num, int nrows,
                             const int64_t *row_ptr, const int32_t *frames, const int32_t *counts, float *G2, float *IP,
                             float *IF)
{
    StSched sc;
    sc.F = F;
    sc.T = T;
    sc.cnt0 = cnt0;
    sc.lastl = lastl;
    sc.cnt_last = cnt_last;
    sc.k = k;
    sc.ev_num = ev_num;
    if (dpl == 8) return run<8>(sc, nrows, row_ptr, frames, counts, G2, IP, IF);
    if (dpl == 4) return run<4>(sc, nrows, row_ptr, frames, counts, G2, IP, IF);
    return 1;
}

// the source lines of the core's barriers met so far; the one to drop in the following runs (0 = none)
extern "C" int mt_stream_emu_sites(int *lines, int cap)
{
    for (int i = 0; i < xpcs::st::g_nsites && i < cap; i++) lines[i] = xpcs::st::g_sites[i];
    return xpcs::st::g_nsites;
}
extern "C" void mt_stream_emu_drop(int line) { xpcs::st::g_drop_line = line; }

// stand-alone driver for the ThreadSanitizer build: reads the arrays from a file written by the test, writes G2 | IP | IF
#ifdef ST_EMU_MAIN
#include <stdio.h>
int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *fp = fopen(argv[1], "rb");
    if (!fp) return 2;
    int32_t hdr[9];
    if (fread(hdr, sizeof(int32_t), 9, fp) != 9) return 2;
    const int dpl = hdr[0], F = hdr[1], T = hdr[2], cnt0 = hdr[3], lastl = hdr[4], cnt_last = hdr[5], k = hdr[6], ev_num = hdr[7], nrows = hdr[8];
    std::vector<int64_t> ptr((size_t)nrows + 1);
    if (fread(ptr.data(), sizeof(int64_t), ptr.size(), fp) != ptr.size()) return 2;
    std::vector<int32_t> fr((size_t)ptr[nrows]), ct((size_t)ptr[nrows]);
    if (fread(fr.data(), sizeof(int32_t), fr.size(), fp) != fr.size()) return 2;
    if (fread(ct.data(), sizeof(int32_t), ct.size(), fp) != ct.size()) return 2;
    fclose(fp);
    std::vector<float> out((size_t)3 * T * nrows, 0.f);
    const int rc = mt_stream_emu(dpl, F, T, cnt0, lastl, cnt_last, k, ev_num, nrows, ptr.data(), fr.data(), ct.data(), out.data(),
                                 out.data() + (size_t)T * nrows, out.data() + (size_t)2 * T * nrows);
    if (rc) return rc;
    fp = fopen(argv[2], "wb");
    fwrite(out.data(), sizeof(float), out.size(), fp);
    fclose(fp);
    return 0;
}
#endif
