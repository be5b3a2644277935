// mt_slicef_host.cpp -- test infrastructure: runs the per-row routines of the CUDA kernel k_multitau_slicef
// (xpcs-eigen_b200/csrc/multitau_slicef_core.h on top of multitau_slice_core.h with XS_CB = 0, the very source nvcc
// compiles for the device) on the CPU, one lane after the other, with the kernel's own split of the work: compat
// phases, pair pieces with one accumulator array each, dense pieces with one slot each, IF / IP ranges, and the
// kernel's order of adding the pieces up.  tests/test_multitau_slicef_core.py compares the result with the oracle.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define XS_NS slf
#define XS_CB 0
#include "../../xpcs-eigen_b200/csrc/multitau_slicef_core.h"

using namespace xpcs::slf;

template <int DPL, bool COMPAT>
static int run(const SlSched &sc, const int *row_len, const uint32_t *frames, const float *values, int len, int ld_factor,
               int ld_cap, int np, int nps, int nd, int nio, int nwarps, float *G2, float *IP, float *IF)
{
    const int T = sc.T, nl = sc.nl, F = sc.F;
    std::vector<uint32_t> frS((size_t)(len + 1) * 32, kSent), lim((size_t)nl * 32, 0u);
    std::vector<float> vlS((size_t)(len + 1) * 32, -1.0e30f);  // (never read past the end of a row)
    std::vector<uint32_t> cntml((size_t)kMlRows * 32, 0u), nlive((size_t)nl * 32, 0u), sbx((size_t)nl * 32, 0u);
    for (int r = 0; r < 32; r++)
        for (int j = 0; j < row_len[r]; j++) {
            frS[(size_t)j * 32 + r] = frames[(size_t)j * 32 + r];
            vlS[(size_t)j * 32 + r] = values[(size_t)j * 32 + r];
        }
    int ld = std::min(nl, ld_cap);
    for (int l = 1; l < ld; l++)
        if ((F >> l) <= ld_factor * std::max(len, 1)) {
            ld = l;
            break;
        }
    const int hsp = std::min(T, sc.cnt0 + DPL * (ld - 1));
    const int h_rows = std::min(T, sc.cnt0 + DPL * (ld_cap - 1));
    if (hsp > h_rows) return -2;
    const int target = ld <= sc.lastl ? dense_target<DPL>(sc, ld, nd) : 0;
    int ndp = 0;
    for (int l = ld; l <= sc.lastl; l++) ndp += dense_pieces(sc, l, target);
    std::vector<float> Hp((size_t)(np + nps) * h_rows * 32, 0.0f), Dp((size_t)std::max(ndp, 1) * DPL * 32, -7.0f);
    for (int lane = 0; lane < 32; lane++) {
        const uint32_t *fr = frS.data() + lane;
        const float *vl = vlS.data() + lane;
        const int n = row_len[lane];
        if (COMPAT) {
            const int chunk = (len + nwarps - 1) / nwarps;
            for (int w = 0; w < nwarps; w++) lane_mlhist(fr, w * chunk, std::min(n, w * chunk + chunk), cntml.data() + lane);
            for (int l = 1; l <= sc.lastl; l++)
                lane_level_base(fr, n, l, ld, F, cntml.data() + lane, nlive.data() + lane, sbx.data() + lane, true);
            nlive[lane] = (uint32_t)n;
            sbx[lane] = (uint32_t)kInfKey;
            for (int l = 0; l < nl; l++) {
                const int Ll = F >> l;
                uint32_t v = l < ld ? (uint32_t)Ll << l : (uint32_t)Ll;
                if (l >= 1 && l <= sc.lastl) v = lane_level_limit(fr, n, l, ld, F, nlive.data() + lane, sbx.data() + lane);
                lim[(size_t)l * 32 + lane] = v;
            }
        } else {
            for (int l = 0; l < nl; l++) {
                const int Ll = F >> l;
                lim[(size_t)l * 32 + lane] = l < ld ? (uint32_t)Ll << l : (uint32_t)Ll;
            }
        }
        const double total = lanef_total(vl, n);
        for (int w = 0; w < np + nps; w++) {
            const int cut = nps > 0 ? n - (n >> 2) : n;
            int piece = w, ia = 0, ib = cut, istep = np;
            if (piece >= np) {
                piece -= np;
                ia = cut;
                ib = n;
                istep = nps;
            }
            float *H = Hp.data() + (size_t)w * h_rows * 32 + lane;
            if (ld - 1 < sc.lastl) lanef_pairs<DPL, true>(fr, vl, ia, ib, piece, istep, ld, sc, lim.data() + lane, H);
            else lanef_pairs<DPL, false>(fr, vl, ia, ib, piece, istep, ld, sc, lim.data() + lane, H);
        }
        for (int part = 0; part < nio; part++) {
            const int ta = (int)((int64_t)T * part / nio), tb = (int)((int64_t)T * (part + 1) / nio);
            lanef_if<DPL>(fr, vl, n, total, sc, ta, tb, IF + (size_t)ta * 32 + lane, 32);
            lanef_ip<DPL>(fr, vl, n, total, sc, ta, tb, IP + (size_t)ta * 32 + lane, 32);
        }
        {
            int piece = 0;
            for (int l = ld; l <= sc.lastl; l++) {
                const int Ll = F >> l;
                const int npieces = dense_pieces(sc, l, target);
                for (int k = 0; k < npieces; k++, piece++) {
                    const int tb = k * target;
                    const int te = k == npieces - 1 ? Ll : tb + target;
                    double acc[DPL];
                    for (int d = 0; d < DPL; d++) acc[d] = 0.0;
                    lanef_dense<DPL>(fr, vl, n, l, tb, te, (int)lim[(size_t)l * 32 + lane], acc);
                    for (int d = 0; d < DPL; d++) Dp[((size_t)piece * DPL + d) * 32 + lane] = (float)acc[d];
                }
            }
        }
        for (int ti = 0; ti < T; ti++) {
            double num = 0.0;
            if (ti < hsp) {
                for (int w = 0; w < np + nps; w++) num += (double)Hp[((size_t)w * h_rows + ti) * 32 + lane];
            } else {
                const int q = ti - sc.cnt0, l = 1 + q / DPL, d = q % DPL;
                int base = 0;
                for (int j = ld; j < l; j++) base += dense_pieces(sc, j, target);
                const int npieces = dense_pieces(sc, l, target);
                for (int k = 0; k < npieces; k++) num += (double)Dp[((size_t)(base + k) * DPL + d) * 32 + lane];
            }
            G2[(size_t)ti * 32 + lane] = g2f_value<DPL>(num, ti, sc);
        }
    }
    return 0;
}

extern "C" int mt_slicef_host(int dpl, int compat, int F, int nl, int T, int cnt0, int lastl, int cnt_last,
                              const int *row_len, const uint32_t *frames, const float *values, int len, int ld_factor,
                              int ld_cap, int np, int nps, int nd, int nio, int nwarps, float *G2, float *IP, float *IF)
{
    SlSched sc{F, nl, T, cnt0, lastl, cnt_last};
    if (dpl == 8) return compat ? run<8, true>(sc, row_len, frames, values, len, ld_factor, ld_cap, np, nps, nd, nio, nwarps, G2, IP, IF)
                                : run<8, false>(sc, row_len, frames, values, len, ld_factor, ld_cap, np, nps, nd, nio, nwarps, G2, IP, IF);
    if (dpl == 4) return compat ? run<4, true>(sc, row_len, frames, values, len, ld_factor, ld_cap, np, nps, nd, nio, nwarps, G2, IP, IF)
                                : run<4, false>(sc, row_len, frames, values, len, ld_factor, ld_cap, np, nps, nd, nio, nwarps, G2, IP, IF);
    return -1;
}
