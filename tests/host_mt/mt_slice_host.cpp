// mt_slice_host.cpp -- test infrastructure: runs the per-row routines of the CUDA kernel k_multitau_slice
// (xpcs-eigen_b200/csrc/multitau_slice_core.h, the very source nvcc compiles for the device) on the CPU,
// one lane after the other, with the kernel's own split of the work (pair pieces, IF / IP ranges, dense
// pieces or 8-bit bin arrays, compat phases).  tests/test_multitau_slice_core.py compares the result bit
// for bit with the oracle.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../xpcs-eigen_b200/csrc/multitau_slice_core.h"

using namespace xpcs::sl;

template <int DPL, bool COMPAT>
static int run(const SlSched &sc, const int *row_len, const uint32_t *words, int len, int ld_factor, int ld_cap,
               int np, int nd, int nio, int nwarps, int bins_rows, float *G2, float *IP, float *IF)
{
    const int T = sc.T, nl = sc.nl, F = sc.F;
    std::vector<uint32_t> evS((size_t)(len + 1) * 32, kSent), H((size_t)T * 32, 0u), lim((size_t)nl * 32, 0u);
    std::vector<uint32_t> cntml((size_t)kMlRows * 32, 0u), nlive((size_t)nl * 32, 0u), sbx((size_t)nl * 32, 0u);
    uint32_t tot[32];
    bool small = true;
    for (int r = 0; r < 32; r++) {
        tot[r] = 0;
        for (int j = 0; j < row_len[r]; j++) {
            evS[(size_t)j * 32 + r] = words[(size_t)j * 32 + r];
            tot[r] += words[(size_t)j * 32 + r] & kCMask;
        }
        if (tot[r] >= 65536u) return 1;
        if (tot[r] > 255u) small = false;
    }
    int ld = std::min(nl, ld_cap);
    for (int l = 1; l < ld; l++)
        if ((F >> l) <= ld_factor * std::max(len, 1)) {
            ld = l;
            break;
        }
    const bool use8 = bins_rows > 0 && ld <= sc.lastl && (F >> ld) <= bins_rows && small;
    std::vector<uint8_t> Bb[3];
    for (int t = 0; t < 3; t++) Bb[t].assign((size_t)((bins_rows >> t) + 1) * 32, 0xee);
    const int hsp = std::min(T, sc.cnt0 + DPL * (ld - 1));
    for (int lane = 0; lane < 32; lane++) {
        const uint32_t *ev = evS.data() + lane;
        const int n = row_len[lane];
        int smin0 = kInfKey;
        if (COMPAT) {
            const int chunk = (len + nwarps - 1) / nwarps;
            for (int w = 0; w < nwarps; w++) lane_mlhist(ev, w * chunk, std::min(n, w * chunk + chunk), cntml.data() + lane);
            for (int l = 1; l <= sc.lastl; l++)
                lane_level_base(ev, n, l, ld, F, cntml.data() + lane, nlive.data() + lane, sbx.data() + lane, !use8);
            nlive[lane] = (uint32_t)n;
            sbx[lane] = (uint32_t)kInfKey;
            for (int l = 0; l < nl; l++) {
                const int Ll = F >> l;
                uint32_t v = l < ld ? (uint32_t)Ll << l : (uint32_t)Ll;
                if (l >= 1 && l <= sc.lastl && (!use8 || l <= ld)) v = lane_level_limit(ev, n, l, ld, F, nlive.data() + lane, sbx.data() + lane);
                lim[(size_t)l * 32 + lane] = v;
            }
            for (int j = 1; j <= std::min(ld, sc.lastl); j++) smin0 = std::min(smin0, (int)sbx[(size_t)j * 32 + lane]);
        } else {
            for (int l = 0; l < nl; l++) {
                const int Ll = F >> l;
                lim[(size_t)l * 32 + lane] = l < ld ? (uint32_t)Ll << l : (uint32_t)Ll;
            }
        }
        const int nps = np / 2;  // small pieces over the last quarter of the events, as in the kernel (0: none)
        for (int w = 0; w < np + nps; w++) {
            const int cut = nps > 0 ? n - (n >> 2) : n;
            int piece = w, ia = 0, ib = cut, istep = np;
            if (piece >= np) {
                piece -= np;
                ia = cut;
                ib = n;
                istep = nps;
            }
            if (ld - 1 < sc.lastl) lane_pairs<DPL, true>(ev, ia, ib, piece, istep, ld, sc, lim.data() + lane, H.data() + lane);
            else lane_pairs<DPL, false>(ev, ia, ib, piece, istep, ld, sc, lim.data() + lane, H.data() + lane);
        }
        for (int part = 0; part < nio; part++) {
            const int ta = (int)((int64_t)T * part / nio), tb = (int)((int64_t)T * (part + 1) / nio);
            lane_if<DPL>(ev, n, tot[lane], sc, ta, tb, IF + (size_t)ta * 32 + lane, 32);
            lane_ip<DPL>(ev, n, tot[lane], sc, ta, tb, IP + (size_t)ta * 32 + lane, 32);
        }
        if (ld <= sc.lastl && use8) {
            for (int t = 0; t < 3; t++) {
                if (ld + t > sc.lastl) continue;
                for (int k = 0; k < (F >> (ld + t)); k++) Bb[t][(size_t)k * 32 + lane] = 0;
                lane_dense8<DPL, COMPAT>(t, ev, n, ld, sc, Bb[t].data() + lane, lim.data() + lane, nlive.data() + lane, smin0,
                                         G2 + lane, 32);
            }
        } else if (ld <= sc.lastl) {
            constexpr int W = 2 * DPL + 1;
            int bins = 0;
            for (int l = ld; l <= sc.lastl; l++) bins += F >> l;
            int target = ((bins + nd - 1) / nd + W - 1) / W * W;
            if (target < W) target = W;
            for (int l = ld; l <= sc.lastl; l++) {
                const int Ll = F >> l;
                const int cnt = level_count<DPL>(sc, l);
                const int npieces = std::max(1, (Ll + target - 1) / target);
                for (int k = 0; k < npieces; k++) {
                    const int tb = k * target;
                    const int te = k == npieces - 1 ? Ll : tb + target;
                    uint32_t acc[DPL];
                    for (int d = 0; d < DPL; d++) acc[d] = 0u;
                    lane_dense<DPL>(ev, n, l, tb, te, (int)lim[(size_t)l * 32 + lane], acc);
                    for (int d = 0; d < DPL; d++)
                        if (d < cnt) H[(size_t)(sc.cnt0 + (l - 1) * DPL + d) * 32 + lane] += acc[d];
                }
            }
        }
        for (int ti = 0; ti < (use8 ? hsp : T); ti++) G2[(size_t)ti * 32 + lane] = g2_value<DPL>(H[(size_t)ti * 32 + lane], ti, sc);
    }
    return use8 ? 2 : 0;
}

// returns 1: slice flagged for the fallback kernel, 0: done with the on-the-fly dense walk, 2: done with 8-bit bins
extern "C" int mt_slice_host(int dpl, int compat, int F, int nl, int T, int cnt0, int lastl, int cnt_last,
                             const int *row_len, const uint32_t *words, int len, int ld_factor, int ld_cap, int np, int nd, int nio,
                             int nwarps, int bins_rows, float *G2, float *IP, float *IF)
{
    SlSched sc{F, nl, T, cnt0, lastl, cnt_last};
    if (dpl == 8) return compat ? run<8, true>(sc, row_len, words, len, ld_factor, ld_cap, np, nd, nio, nwarps, bins_rows, G2, IP, IF)
                                : run<8, false>(sc, row_len, words, len, ld_factor, ld_cap, np, nd, nio, nwarps, bins_rows, G2, IP, IF);
    if (dpl == 4) return compat ? run<4, true>(sc, row_len, words, len, ld_factor, ld_cap, np, nd, nio, nwarps, bins_rows, G2, IP, IF)
                                : run<4, false>(sc, row_len, words, len, ld_factor, ld_cap, np, nd, nio, nwarps, bins_rows, G2, IP, IF);
    return -1;
}
