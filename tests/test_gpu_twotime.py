"""Two-time correlation (tcgen05 kernel, through the C-ABI) against the CPU oracle on seeded
inputs and against the golden fixtures written by the unmodified reference binary.
Tolerance: 1e-5 relative (BASELINE.json north_star) on C, g2full, g2partials; sg exact for
integer photon counts."""
import numpy as np
import pytest

import golden_util as G
from conftest import make_case

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _close(a, b, what, rtol=RTOL, atol=0.0):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    assert a.size == b.size, what
    bad = np.abs(a - b) > atol + rtol * np.abs(b)
    assert not bad.any(), "%s: %d of %d beyond rtol %g (worst %.3g at %d: %r vs %r)" % (
        what, bad.sum(), a.size, rtol, np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)), np.argmax(bad),
        a[np.argmax(bad)], b[np.argmax(bad)])


def _gpu_twotime(pkg, dq, sq, F, off, idx, val, qbins, wsize, average=False, method="symmetric", **kw):
    c = pkg.Correlator(dq, sq, F, **kw)
    c.push_sparse(idx, val, off)
    c.finish_ingest(want=False)
    out = {q: c.twotime(q, wsize, method=method, average=average) for q in qbins}
    launches = c.kernel_report()
    c.close()
    return out, launches


@pytest.mark.parametrize("name", G.names("twotime"))
def test_twotime_matches_reference_fixture(pkg, name):
    c = G.Case(name)
    qbins = [int(q) for q in c.inp["qbins"]]
    avg = str(c.inp["filt"]) == "Average"
    out, launches = _gpu_twotime(pkg, c.dq, c.sq, c.F, c.inp["off"], c.inp["idx"], c.inp["val"], qbins,
                                 int(c.inp["wsize"]), average=avg, method=c.method, static_window=c.swindow)
    assert launches.get("k_twotime_gemm", (0, 0))[1] == len(qbins)
    sg_row = 0
    for b, q in enumerate(qbins):
        r = out[q]
        if c.method == "staticmap":  # one sg row per static partition of the processed bins, in map order
            n = r["sg"].shape[0]
            assert G.n_diff(r["sg"], c.ref["sg"][sg_row: sg_row + n]) == 0, "sg of bin %d" % q
            sg_row += n
        else:
            assert G.n_diff(r["sg"], c.ref["sg"][b]) == 0, "sg of bin %d" % q
        _close(r["C"], c.ref["C2T_all/g2_%05d" % q], "C bin %d" % q)
        assert not np.tril(r["C"], -1).any(), "lower triangle must stay zero"
        _close(r["g2full"], c.ref["g2full"][:, b], "g2full bin %d" % q)
        _close(r["g2partials"], c.ref["g2partials"][:, :, b], "g2partials bin %d" % q)


@pytest.mark.parametrize("h,w,F,occ,seed,wsize", [
    (24, 24, 130, 0.08, 1, 10),      # F just over one 128 tile: partial tiles on both axes
    (40, 40, 300, 0.05, 2, 25),      # 3x3 tiles, several K blocks
    (64, 64, 517, 0.02, 3, 50),      # odd F
])
def test_twotime_matches_oracle(pkg, oracle, h, w, F, occ, seed, wsize):
    dq, sq, off, idx, val = make_case(pkg, h, w, F, occ, seed, n_dynamic=3, static_per_dynamic=2)
    qm = oracle.QMap(dq, sq)
    fo = oracle.sparse_filter(qm, F, off, idx, val, swindow=max(1, F // 10))
    qbins = [1, 2, 3]
    ref = oracle.twotime(qm, F, fo.rows, qbins, wsize, method="symmetric", average=False)
    out, _ = _gpu_twotime(pkg, dq, sq, F, off, idx, val, qbins, wsize)
    for b, q in enumerate(ref["bins"]):
        assert G.n_diff(out[q]["sg"], ref["sg"][b]) == 0
        _close(out[q]["C"], ref["C"][q], "C bin %d" % q)
        _close(out[q]["g2full"], ref["g2full"][:, b], "g2full bin %d" % q)
        _close(out[q]["g2partials"], ref["g2partials"][:, :, b], "g2partials bin %d" % q)


def test_twotime_float_rows_three_pass(pkg, oracle):
    """Flat-fielded (float) rows: fp16 hi/lo split, three tensor-core passes."""
    h, w, F = 32, 32, 260
    dq, sq, off, idx, val = make_case(pkg, h, w, F, 0.06, 5, n_dynamic=2, static_per_dynamic=2)
    flat = pkg.synth.flatfield(h * w)
    qm = oracle.QMap(dq, sq)
    fo = oracle.sparse_filter(qm, F, off, idx, val, flat=flat, swindow=26)
    ref = oracle.twotime(qm, F, fo.rows, [1, 2], 20, method="symmetric", average=False)
    out, _ = _gpu_twotime(pkg, dq, sq, F, off, idx, val, [1, 2], 20, flatfield=flat)
    for b, q in enumerate(ref["bins"]):
        _close(out[q]["sg"], ref["sg"][b], "sg bin %d" % q, 1e-6)
        _close(out[q]["C"], ref["C"][q], "C bin %d" % q)
        _close(out[q]["g2full"], ref["g2full"][:, b], "g2full bin %d" % q)


@pytest.mark.parametrize("average", [False, True])
def test_twotime_staticmap_matches_oracle(pkg, oracle, average):
    """StaticMap smoothing (SmoothingStaticMap / ComputeSGStaticMap, corr.cpp:433-494, :1228-1305): every pixel is
    divided by the sg of its own static partition; 5 static partitions per dynamic bin, several K blocks."""
    h, w, F, wsize = 48, 48, 300, 25
    dq, sq, off, idx, val = make_case(pkg, h, w, F, 0.05, 7, n_dynamic=2, static_per_dynamic=5)
    qm = oracle.QMap(dq, sq)
    fo = oracle.sparse_filter(qm, F, off, idx, val, swindow=30)
    ref = oracle.twotime(qm, F, fo.rows, [1, 2], wsize, method="staticmap", average=average)
    out, _ = _gpu_twotime(pkg, dq, sq, F, off, idx, val, [1, 2], wsize, average=average, method="staticmap")
    row = 0
    for b, q in enumerate(ref["bins"]):
        n = out[q]["sg"].shape[0]
        assert n == 5
        assert G.n_diff(out[q]["sg"], ref["sg"][row: row + n]) == 0
        row += n
        _close(out[q]["C"], ref["C"][q], "C bin %d" % q)
        _close(out[q]["g2full"], ref["g2full"][:, b], "g2full bin %d" % q)
        _close(out[q]["g2partials"], ref["g2partials"][:, :, b], "g2partials bin %d" % q)


def test_twotime_counts_above_2048_take_the_split_operand(pkg, oracle):
    """fp16 holds integers exactly up to 2048 only: packed counts above that (the store admits 4095) must go
    through the hi + lo operand (three passes) instead of being rounded."""
    h, w, F, wsize = 24, 24, 200, 20
    dq, sq, off, idx, val = make_case(pkg, h, w, F, 0.06, 11, n_dynamic=2, static_per_dynamic=2)
    val = val.copy()
    val[::7] = 4001   # odd, above 2048: not representable in fp16
    val[::11] = 2049
    qm = oracle.QMap(dq, sq)
    fo = oracle.sparse_filter(qm, F, off, idx, val, swindow=20)
    ref = oracle.twotime(qm, F, fo.rows, [1, 2], wsize, method="symmetric", average=False)
    c = pkg.Correlator(dq, sq, F)
    c.push_sparse(idx, val, off)
    c.finish_ingest(want=False)
    assert c.info().value_kind == 0
    for b, q in enumerate(ref["bins"]):
        r = c.twotime(q, wsize)
        assert G.n_diff(r["sg"], ref["sg"][b]) == 0
        _close(r["C"], ref["C"][q], "C bin %d" % q)
        _close(r["g2full"], ref["g2full"][:, b], "g2full bin %d" % q)
    c.close()
