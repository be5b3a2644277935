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
                                 int(c.inp["wsize"]), average=avg, static_window=c.swindow)
    assert launches.get("k_twotime_gemm", (0, 0))[1] == len(qbins)
    for b, q in enumerate(qbins):
        r = out[q]
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
