"""The warp-per-row multi-tau kernel (csrc/multitau_warp.cu) against the oracle and against the
lane-per-row kernel (csrc/multitau.cu), through the C-ABI.  Integer counts: bit-exact at every
level, with and without the reference's stale-tail behaviour.  Covers the
shapes of the bench workloads (100 k and 1 M frames at ~100 events per row), rows of every
density inside one slice, and the two fallbacks (rows longer than the kernel's shared-memory
budget, rows whose counts sum to 2^16 or more)."""
import numpy as np
import pytest

from conftest import make_case
from test_gpu_parity import assert_exact, run_oracle

pytestmark = pytest.mark.gpu


def frames_to_stream(P, F, rows_f, rows_c):
    """pixel-major (frames, counts) lists -> frame-major (off, idx, val)."""
    pix = np.concatenate([np.full(len(f), p, np.int32) for p, f in enumerate(rows_f)])
    fr = np.concatenate([np.asarray(f, np.int64) for f in rows_f])
    cnt = np.concatenate([np.asarray(c, np.int64) for c in rows_c])
    order = np.lexsort((pix, fr))
    pix, fr, cnt = pix[order], fr[order], cnt[order]
    off = np.zeros(F + 1, np.int64)
    np.add.at(off, fr + 1, 1)
    return np.cumsum(off), pix.astype(np.int32), cnt.astype(np.int16)


def gpu_multitau(pkg, dq, sq, F, off, idx, val, **kw):
    c = pkg.Correlator(dq, sq, F, **kw)
    c.push_sparse(idx, val, off)
    c.finish_ingest(want=False)
    G = c.multitau()
    g2, _ = c.normalize()
    info = c.info()
    fb = c.multitau_fallback_slices()
    n_slices = (info.n_rows + 31) // 32
    c.close()
    return G, g2, info, fb, n_slices


def check(pkg, oracle, dq, sq, F, off, idx, val, dpl=8, compat=True, also_lane=True, expect_fallback=None):
    G, g2, info, fb, n_slices = gpu_multitau(pkg, dq, sq, F, off, idx, val, dpl=dpl, compat=compat)
    assert info.value_kind == 0
    assert fb >= 0, "the warp-per-row kernel did not run"
    if expect_fallback is not None:
        assert expect_fallback(fb, n_slices), "fallback slices: %d of %d" % (fb, n_slices)
    _, rG, rg2, _ = run_oracle(oracle, dq, sq, F, off, idx, val, dpl=dpl, compat=compat)
    for k, name in enumerate(("G2", "IP", "IF")):
        assert_exact(G[k], rG[k], name + " (warp kernel vs oracle)")
    assert_exact(g2, rg2, "norm-0-g2")
    if also_lane:
        GL, _, _, fbl, _ = gpu_multitau(pkg, dq, sq, F, off, idx, val, dpl=dpl, compat=compat, lane_multitau=True)
        assert fbl == -1
        for k, name in enumerate(("G2", "IP", "IF")):
            assert_exact(G[k], GL[k], name + " (warp kernel vs lane kernel)")


@pytest.mark.parametrize("h,w,F,occ,seed,dpl", [
    (64, 64, 100000, 0.001, 11, 8),     # bench workload c3: 14 levels, ~100 events per row
    (48, 48, 20000, 0.02, 12, 8),       # ~400 events per row
    (32, 32, 1000000, 0.0001, 13, 8),   # c5 shape: 17 levels, T = 142
    (40, 40, 10000, 0.01, 14, 4),
    (32, 32, 1500, 0.2, 15, 8),         # dense rows: every level but 0 on the bin arrays
    (32, 32, 500, 0.9, 16, 4),
])
@pytest.mark.parametrize("compat", [True, False])
def test_bench_shapes(pkg, oracle, h, w, F, occ, seed, dpl, compat):
    dq, sq, off, idx, val = make_case(pkg, h, w, F, occ, seed)
    check(pkg, oracle, dq, sq, F, off, idx, val, dpl=dpl, compat=compat, also_lane=compat,
          expect_fallback=lambda fb, ns: fb == 0)


@pytest.mark.parametrize("F,dpl,seed", [(512, 8, 1), (4096, 8, 2), (2500, 4, 3), (33, 8, 4), (100000, 8, 5)])
def test_mixed_density_slices(pkg, oracle, F, dpl, seed):
    """Rows of very different density next to each other in one slice, clustered rows (the
    stale-tail regime), single-event and empty rows."""
    rng = np.random.default_rng(seed)
    h, w = 16, 16
    P = h * w
    dq = np.ones((h, w), np.int32)
    sq = (1 + (np.arange(P) // 64)).astype(np.int32).reshape(h, w)
    dens = [0.002, 0.01, 0.03, 0.08, 0.2, 0.5, 0.9, 1.0, "cluster", 0.0, "one"]
    rows_f, rows_c = [], []
    for p in range(P):
        d = dens[p % len(dens)]
        if d == "cluster":
            base = int(rng.integers(0, max(F - 40, 1)))
            f = np.unique(np.concatenate([rng.integers(0, F, 3), base + rng.integers(0, 40, 25)]))
            f = f[f < F]
        elif d == "one":
            f = np.array([int(rng.integers(0, F))])
        else:
            scale = min(1.0, 3000.0 / F) * (0.1 if (p // 32) % 2 == 0 else 1.0)  # even slices stay short
            f = np.nonzero(rng.random(F) < d * scale)[0]
        rows_f.append(f)
        rows_c.append(1 + rng.poisson(0.3, f.size))
    if F == 512:  # the hand example of SURVEY.md A.4
        rows_f[0] = np.array(list(range(32)) + [400, 440])
        rows_c[0] = np.ones(34, np.int64)
    off, idx, val = frames_to_stream(P, F, rows_f, rows_c)
    check(pkg, oracle, dq, sq, F, off, idx, val, dpl=dpl, compat=True, expect_fallback=lambda fb, ns: fb <= ns // 2)
    check(pkg, oracle, dq, sq, F, off, idx, val, dpl=dpl, compat=False, also_lane=False)


def test_heavy_and_long_rows_fall_back(pkg, oracle):
    """One pixel lit in every frame (longer than the warp kernel's shared-memory budget) and one
    whose counts sum beyond 2^16 (32-bit numerators could overflow): their slices are redone by
    the lane-per-row kernel, all other slices stay on the warp kernel; results are exact."""
    F = 6000
    h, w = 16, 16
    P = h * w
    dq = np.ones((h, w), np.int32)
    sq = (1 + (np.arange(P) // 64)).astype(np.int32).reshape(h, w)
    rng = np.random.default_rng(9)
    rows_f = [np.nonzero(rng.random(F) < 0.01)[0] for _ in range(P)]
    rows_c = [1 + rng.poisson(0.2, f.size) for f in rows_f]
    rows_f[5] = np.arange(F)                          # long row
    rows_c[5] = 1 + rng.poisson(0.2, F)
    rows_f[100] = np.sort(rng.choice(F, 40, replace=False))   # heavy row: 40 * 2000 = 80000 >= 2^16
    rows_c[100] = np.full(40, 2000)
    off, idx, val = frames_to_stream(P, F, rows_f, rows_c)
    check(pkg, oracle, dq, sq, F, off, idx, val, compat=True, expect_fallback=lambda fb, ns: fb == 2 and ns == 8)
