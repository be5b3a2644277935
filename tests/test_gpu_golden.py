"""CUDA path (through the C-ABI) against the golden fixtures written by the unmodified
reference binary (tests/golden/make_golden.py).  Integer photon-count inputs: bit-exact on
every dataset.  Float inputs (flat-field, averaging, frame-sum normalisation, dense source):
within the 1e-5 relative tolerance of BASELINE.json's north_star."""
import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _close(a, b, what, rtol=RTOL):
    err, nan_mismatch = G.rel_err(a, b)
    assert nan_mismatch == 0, "%s: NaN pattern differs" % what
    assert err <= rtol, "%s: worst relative error %.3g > %g" % (what, err, rtol)


def _run(pkg, c):
    kw = dict(dpl=c.dpl, flatfield=c.flat, stride=c.stride, avg=c.avg, static_window=c.swindow,
              normalize_by_framesum=bool(c.norm), compat=True, late_window=c.late_window)
    if c.kind == "dense" and "thresh" in c.inp:
        kw.update(lld=float(c.inp["thresh"][0]), sigma=float(c.inp["thresh"][1]))
    cor = pkg.Correlator(c.dq, c.sq, c.F, **kw)
    dark = None
    if c.kind == "sparse":
        cor.push_sparse(c.inp["idx"], c.inp["val"], c.inp["off"])
    else:
        fr = c.inp["frames"]
        if c.darks:
            cor.set_dark(fr[: c.darks])
            dark = cor.get_dark()
        cor.push_dense(fr[c.darks:])
    sums = cor.finish_ingest()
    Gs = cor.multitau()
    g2, se = cor.normalize()
    kind = cor.info().value_kind
    cor.close()
    return sums, Gs, g2, se, kind, dark


@pytest.mark.parametrize("name", G.names("sparse") + G.names("dense"))
def test_cuda_matches_reference_fixture(pkg, name):
    c = G.Case(name)
    sums, Gs, g2, se, kind, dark = _run(pkg, c)
    exact = kind == 0  # packed integer store
    chk = (lambda a, b, w: (G.n_diff(a, b) == 0) or pytest.fail("%s: %d entries differ" % (w, G.n_diff(a, b)))) \
        if exact else _close
    for k, nm in enumerate(("G2", "IP", "IF")):
        chk(Gs[k], c.ref[nm], nm)
    chk(sums["frame_sum"], c.ref["frameSum"], "frameSum")
    chk(sums["pixel_sum"], c.ref["pixelSum"], "pixelSum")
    chk(sums["part_total"], c.ref["partition-mean-total"], "partition-mean-total")
    chk(sums["part_partial"], c.ref["partition-mean-partial"], "partition-mean-partial")
    chk(g2, c.ref["norm-0-g2"], "norm-0-g2")
    _close(se, c.ref["norm-0-stderr"], "norm-0-stderr")
    if dark is not None:
        _close(dark[0], c.ref["DarkAvg"], "DarkAvg", 1e-12)
        _close(dark[1], c.ref["DarkStd"], "DarkStd", 1e-12)


def test_integer_fixtures_use_the_exact_path(pkg):
    for name in ("sparse_int_24x24", "sparse_staletail_32x32", "staletail_hand_example"):
        assert _run(pkg, G.Case(name))[4] == 0
