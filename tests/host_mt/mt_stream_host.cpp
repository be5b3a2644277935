// mt_stream_host.cpp -- test infrastructure: runs the per-row routines of the online multi-tau kernel
// k_stream_chunk (xpcs-eigen_b200/csrc/multitau_stream_core.h, the very source nvcc compiles for the device) on
// the CPU: the frames of every row are cut into chunks of 2^k frames, each chunk goes through row_chunk() with the
// 32 lanes of every phase run one after the other, and row_result() forms the outputs from the state at the end.
// tests/test_multitau_stream_core.py compares them bit for bit with the oracle (exact maths, no stale-tail flag).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../xpcs-eigen_b200/csrc/multitau_stream_core.h"

using namespace xpcs::st;

template <int DPL>
static int run(const StSched &sc, int nrows, const int64_t *row_ptr, const int32_t *frames, const int32_t *counts,
               float *G2, float *IP, float *IF)
{
    typedef Layout<DPL> LY;
    const int K = 1 << sc.k;
    const int nchunks = (sc.F + K - 1) / K;
    const int stride = LY::words(sc);
    std::vector<uint32_t> state((size_t)nrows * stride, 0u);
    std::vector<uint32_t> x((size_t)LY::scratch_words(sc) + 4, 0xdeadbeefu);
    uint32_t *xa = x.data();
    while (reinterpret_cast<uintptr_t>(xa) & 15u) xa++;
    std::vector<uint32_t> ev;
    for (int c = 0; c < nchunks; c++) {
        for (int r = 0; r < nrows; r++) {
            ev.clear();
            for (int64_t j = row_ptr[r]; j < row_ptr[r + 1]; j++)
                if ((frames[j] >> sc.k) == c) {
                    if (counts[j] < 0 || counts[j] > 4095) return 2;
                    ev.resize(ev.size() + 32, 0xffffffffu);
                    ev[ev.size() - 32] = ((uint32_t)frames[j] << kCB) | (uint32_t)counts[j];
                }
            row_chunk<DPL>(sc, c, ev.data(), (int)(ev.size() / 32), xa, state.data() + (size_t)r * stride);
        }
    }
    for (int r = 0; r < nrows; r++)
        for (int ti = 0; ti < sc.T; ti++) {
            const size_t o = (size_t)ti * nrows + r;
            row_result<DPL>(sc, state.data() + (size_t)r * stride, ti, G2[o], IP[o], IF[o]);
        }
    return 0;
}

extern "C" int mt_stream_host(int dpl, int F, int T, int cnt0, int lastl, int cnt_last, int k, int ev_num, int nrows,
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

extern "C" int mt_stream_state_words(int dpl, int T, int lastl)
{
    StSched sc{};
    sc.T = T;
    sc.lastl = lastl;
    return dpl == 8 ? Layout<8>::words(sc) : Layout<4>::words(sc);
}
