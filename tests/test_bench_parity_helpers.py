"""The checker inside bench.py's parity block, checked on the CPU: building oracle rows for a SAMPLE of pixels from
frame-major events, and recomputing norm-0-g2 of ONE dynamic bin from the correlator columns of its pixels, must
give exactly what the oracle gives on the whole detector (otherwise a green "parity" in the bench line would say
nothing)."""
import importlib.util
import os

import numpy as np

from conftest import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_sampled_rows_and_single_bin_normalisation_equal_the_full_oracle(pkg, oracle):
    B, O = _bench(), oracle
    h, w, F = 48, 40, 900
    dq, sq, off, idx, val = make_case(pkg, h, w, F, 0.02, 41, n_dynamic=5, static_per_dynamic=3)
    qm = O.QMap(dq, sq)
    fo = O.sparse_filter(qm, F, off, idx, val, swindow=90)
    G2, IP, IF = O.multitau(h * w, F, 8, fo.rows, compat=True)
    g2, _ = O.normalize(qm, G2, IP, IF)
    # (1) sampled rows
    valid = np.flatnonzero((dq.ravel() > 0) & (sq.ravel() > 0))
    samp = np.sort(np.random.default_rng(1).choice(valid, 150, replace=False)).astype(np.int32)
    fr = np.repeat(np.arange(F), np.diff(off))
    m = np.isin(idx, samp)
    rows = B.rows_of_pixels(O, samp, idx[m], fr[m], val[m])
    rG2, rIP, rIF = O.multitau(samp.size, F, 8, rows, compat=True)
    assert np.array_equal(rG2, G2[:, samp]) and np.array_equal(rIP, IP[:, samp]) and np.array_equal(rIF, IF[:, samp])
    assert (rG2 != 0).sum() > 100
    # (2) one dynamic bin from its pixels' columns
    dqf, sqf = dq.ravel(), sq.ravel()
    for q in (1, 3, 5):
        pix = np.flatnonzero((dqf == q) & (sqf > 0)).astype(np.int32)
        sq_local = np.unique(sqf[pix], return_inverse=True)[1].astype(np.int32) + 1
        qm1 = O.QMap(np.ones((1, pix.size), np.int32), sq_local.reshape(1, -1))
        rg2, _ = O.normalize(qm1, np.ascontiguousarray(G2[:, pix]), np.ascontiguousarray(IP[:, pix]),
                             np.ascontiguousarray(IF[:, pix]))
        assert B.G.n_diff_arrays(rg2[:, 0], g2[:, q - 1]) == 0
