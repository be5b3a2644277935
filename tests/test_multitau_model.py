"""The reformulation used by the warp-per-row multi-tau kernel (tests/multitau_model.py) is
bit-identical to the oracle (reference corr.cpp:315-431, incl. the stale-tail quirk) on integer
rows of every density, including the rows the quirk hits."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as O  # noqa: E402
import multitau_model as M  # noqa: E402


def random_rows(rng, P, F, dens):
    rows_f, rows_c = [], []
    for p in range(P):
        d = dens[p % len(dens)]
        if d == "cluster":
            base = rng.integers(0, max(F - 40, 1))
            f = np.unique(np.concatenate([rng.integers(0, F, 3), base + rng.integers(0, 40, 25)]))
            f = f[f < F]
        else:
            f = np.nonzero(rng.random(F) < d)[0]
        c = 1 + rng.poisson(0.3, f.size)
        rows_f.append(f.astype(np.int32))
        rows_c.append(c.astype(np.int64))
    return rows_f, rows_c


def oracle_rows(rows_f, rows_c, F, dpl, compat):
    P = len(rows_f)
    ptr = np.zeros(P + 1, np.int64)
    ptr[1:] = np.cumsum([len(f) for f in rows_f])
    t = np.concatenate(rows_f).astype(np.int32) if ptr[-1] else np.zeros(0, np.int32)
    v = np.concatenate(rows_c).astype(np.float32) if ptr[-1] else np.zeros(0, np.float32)
    return O.multitau(P, F, dpl, O.Rows(ptr, t, v), compat=compat)


@pytest.mark.parametrize("F,dpl,seed", [(512, 8, 1), (1000, 8, 2), (777, 4, 3), (4096, 8, 4), (33, 8, 5), (15, 8, 6),
                                        (2500, 4, 7)])
@pytest.mark.parametrize("compat", [True, False])
def test_model_matches_oracle(F, dpl, seed, compat):
    rng = np.random.default_rng(seed)
    dens = [0.002, 0.01, 0.03, 0.08, 0.2, 0.5, 0.9, 1.0, "cluster", 0.0]
    P = 60
    rows_f, rows_c = random_rows(rng, P, F, dens)
    # the hand example of SURVEY.md A.4 (F=512): unit counts at 0..31, 400, 440
    if F == 512:
        rows_f[0] = np.array(list(range(32)) + [400, 440], np.int32)
        rows_c[0] = np.ones(34, np.int64)
    G2, IP, IF = oracle_rows(rows_f, rows_c, F, dpl, compat)
    lv, tv = O.delay_schedule(F, dpl)
    sched = M.build_sched(lv, tv)
    stats = {}
    for p in range(P):
        out = M.row_multitau_model(rows_f[p].tolist(), rows_c[p].tolist(), F, dpl, sched, compat=compat, stats=stats)
        for ti, (g, a, b) in out.items():
            assert g == G2[ti, p], (p, ti, g, G2[ti, p])
            assert a == IP[ti, p], (p, ti, a, IP[ti, p])
            assert b == IF[ti, p], (p, ti, b, IF[ti, p])
    if compat and F == 512:
        assert stats.get("lost", 0) > 0  # the quirk is exercised


@pytest.mark.parametrize("dpl", [4, 8])
def test_ipif_threshold_histogram_closed_forms(dpl):
    """The first-delay indices phase 1 of k_multitau_warp derives from bit lengths equal the counts over
    the actual thresholds (corr.cpp:403, 414-416), for every frame of every F tried."""
    for F in list(range(2 * dpl + 2, 300)) + [1000, 1023, 1024, 1025, 4097, 10000]:
        lv, tv = O.delay_schedule(F, dpl)
        sched = M.build_sched(lv, tv)
        lvl = lv.astype(np.int64)
        tp = tv.astype(np.int64) >> lvl
        thr_if = tp << lvl
        thr_ip = np.maximum((F >> lvl) - tp, 0) << lvl
        assert np.all(np.diff(thr_if) > 0) and np.all(np.diff(thr_ip) < 0)
        frames = range(F) if F <= 1025 else list(range(0, 300)) + list(range(F - 300, F)) + list(range(0, F, 37))
        for f in frames:
            a, b = M.ipif_first_delay(f, F, dpl, sched)
            assert a == int(np.sum(thr_if <= f)), (F, f)
            assert b == int(np.sum(thr_ip > f)), (F, f)
