"""Parity at a BASELINE configuration's real size: configs[0] ("C1": sparse IMM 512x512, 10 000 frames, ~1 %
occupancy, 8 delays per level, 36 dynamic / 360 static q-bins; 2.6e7 events, 205 k rows, 6.4 k slices, 88 delays).

The unmodified reference binary (oracle/_ref/corr_ref) runs once on the IMM file; the same file then goes
(a) through the host program `corr` (the reference's entry point) and (b) through the C-ABI with HOST buffers and
the library's DEFAULT knobs -- 2.6e7 events is above the 4 Mi-event threshold, so this is the pipelined,
copy-overlapped, 4-chunk ingest and the warp-per-row multi-tau kernel exactly as the bench runs them (no
XPCS_PIPELINE_* environment, no diagnostic flags).  G2 / IP / IF at every level, norm-0-g2, pixelSum, frameSum and
the partition means are integers-through-IEEE-division and must equal the reference bit for bit; norm-0-stderr
within 1e-5 (fp64 sums instead of the reference's fp32 Welford chain)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import golden_util as G

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refdrv  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # norm-0-stderr only; everything else is compared for equality

H = W = 512
FRAMES = 10000
OCC = 0.01
DPL = 8
EXACT = ("G2", "IP", "IF", "norm-0-g2", "pixelSum", "frameSum", "partition-mean-total", "partition-mean-partial",
         "tau", "partition_norm_factor", "timestamp_clock", "timestamp_tick")


@pytest.fixture(scope="module")
def c1(pkg):
    if not refdrv.available():
        pytest.skip("oracle/_ref/corr_ref not built")
    for k in ("XPCS_NO_PIPELINE", "XPCS_PIPELINE_MIN_EVENTS", "XPCS_PIPELINE_CHUNKS", "XPCS_SCATTER_DIRECT"):
        assert k not in os.environ, "config-scale parity runs with the library's default knobs"
    dq, sq = pkg.synth.annular_qmaps(H, W, n_dynamic=36, static_per_dynamic=10, r_min=8.0)
    off, idx, val = pkg.synth.sparse_frames(H * W, FRAMES, OCC, seed=1234)
    d = refdrv.scratch_dir()
    imm = os.path.join(d, "data.imm")
    pkg.synth.write_imm_sparse(imm, H, W, off, idx, val)
    root = os.path.join(d, "case.h5dir")
    refdrv.write_config(root, dq, sq, FRAMES, imm, dpl=DPL)
    info = refdrv.run(root, imm, g2out=True, threads=len(os.sched_getaffinity(0)), cwd=d)
    ref = refdrv.listing(root, "/exchange")
    shutil.rmtree(root, ignore_errors=True)
    yield dict(dir=d, imm=imm, dq=dq, sq=sq, off=off, idx=idx, val=val, ref=ref, ref_info=info)
    shutil.rmtree(d, ignore_errors=True)


def _compare(res, ref, where):
    assert sorted(res) == sorted(ref), "%s: result dataset names differ from the reference's" % where
    for k, want in ref.items():
        got = res[k]
        assert got.shape == want.shape and got.dtype == want.dtype, "%s %s: %s %s vs %s %s" % (
            where, k, got.shape, got.dtype, want.shape, want.dtype)
        if k in EXACT:
            assert G.n_diff(got, want) == 0, "%s %s: %d entries differ from the reference" % (where, k, G.n_diff(got, want))
        else:
            err, nanmis = G.rel_err(got, want)
            assert nanmis == 0 and err <= RTOL, "%s %s: worst relative error %.3g" % (where, k, err)


def test_c1_corr_program_equals_reference_file(pkg, c1):
    """`corr config.hdf5 --imm data.imm --g2out` against what corr_ref wrote for the same file."""
    corr = os.path.join(os.path.dirname(pkg.cabi.LIB_PATH), "corr")
    cfg = os.path.join(c1["dir"], "config.hdf5")
    f = pkg.h5lite.File()
    for path, value in refdrv.config_items(c1["dq"], c1["sq"], FRAMES, c1["imm"], dpl=DPL)[0]:
        f.put(path, value)
    f.save(cfg)
    f.close()
    p = subprocess.run([corr, cfg, "--g2out"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout[-2000:]
    g = pkg.h5lite.File(cfg)
    res = g.walk("/exchange")
    g.close()
    os.remove(cfg)
    _compare(res, c1["ref"], "corr")
    # the stale-tail quirk is live at this size (level 9 of C1): the exact sums differ from the reference
    assert c1["ref"]["G2"].shape == (88, H * W)


def test_c1_cabi_default_pipeline_equals_reference(pkg, c1):
    """The C-ABI with host buffers and default knobs (pipelined ingest, warp-per-row multi-tau)."""
    c = pkg.Correlator(c1["dq"], c1["sq"], FRAMES, dpl=DPL, compat=True, device=0)
    c.push_sparse(c1["idx"], c1["val"], c1["off"])
    sums = c.finish_ingest()
    info = c.info()
    G2, IP, IF = c.multitau()
    fallback = c.multitau_fallback_slices()
    g2, se = c.normalize()
    rep = c.kernel_report()
    c.close()
    assert info.value_kind == 0 and info.events_stored > 2.0e7
    assert rep.get("k_concat", (0, 0))[1] >= 1, "the pipelined (chunked) ingest did not run: %s" % sorted(rep)
    assert rep.get("k_multitau_slice", (0, 0))[1] >= 1 and fallback == 0, "the short-row kernel did not take C1: %s" % sorted(rep)
    ref = c1["ref"]
    for name, got in (("G2", G2), ("IP", IP), ("IF", IF), ("norm-0-g2", g2), ("pixelSum", sums["pixel_sum"]),
                      ("frameSum", sums["frame_sum"]), ("partition-mean-partial", sums["part_partial"]),
                      ("partition-mean-total", sums["part_total"].reshape(1, -1))):
        assert G.n_diff(got, ref[name]) == 0, "%s: %d entries differ from the reference" % (name, G.n_diff(got, ref[name]))
    err, nanmis = G.rel_err(se, ref["norm-0-stderr"])
    assert nanmis == 0 and err <= RTOL
    # without the compat flag the kernel returns the exact sums, which the reference misses at its deepest level
    c = pkg.Correlator(c1["dq"], c1["sq"], FRAMES, dpl=DPL, compat=False, device=0)
    c.push_sparse(c1["idx"], c1["val"], c1["off"])
    c.finish_ingest(want=False)
    G2x, IPx, IFx = c.multitau()
    c.close()
    assert np.array_equal(IPx, ref["IP"]) and np.array_equal(IFx, ref["IF"])
    diff = np.nonzero(G2x != ref["G2"])
    assert diff[0].size > 0, "expected the reference's dropped pairs (SURVEY A.4) to show at C1"
    assert (G2x[diff] > ref["G2"][diff]).all(), "the reference can only miss pairs"
