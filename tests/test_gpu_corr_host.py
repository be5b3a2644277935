"""End to end through the host program: `corr config.hdf5 --imm data.imm [--g2out] [--darkout] [--ufxc | --rigaku | --hdf5]`
(the reference's entry point) on the inputs of every golden fixture (IMM sparse / dense, UFXC and Rigaku event words, an HDF5 frame stack), results read back from the
configuration HDF5 file and compared dataset by dataset -- name, shape, dtype, values -- with
what the unmodified reference binary wrote (tests/golden/make_golden.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_util as G

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refdrv  # noqa: E402  (config key list shared with the reference runs)

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _run_corr(pkg, c, tmp_path, extra=(), stack_storage=None, corr_path=None, env=None, cfg=None):
    corr = corr_path or os.path.join(os.path.dirname(pkg.cabi.LIB_PATH), "corr")
    imm = str(tmp_path / "data.imm")
    h, w = c.dq.shape
    kw = dict(dpl=c.dpl, stride=c.stride, avg=c.avg, static_window=c.swindow, flatfield=c.flat,
              normalize_by_framesum=bool(c.norm))
    if c.kind == "dense":
        pkg.synth.write_imm_dense(imm, h, w, c.inp["frames"])
        kw.update(darks=c.darks)
        if "thresh" in c.inp:
            kw.update(lld=float(c.inp["thresh"][0]), sigma=float(c.inp["thresh"][1]))
    elif c.fmt == "hdf5":  # a frame stack /entry/data/data (io/hdf5.cpp), read through --hdf5
        st = pkg.h5lite.File()
        st.put("/entry/data/data", c.inp["stack"])
        if stack_storage:
            st.set_storage("/entry/data/data", **stack_storage)
        st.save(imm)
        st.close()
        kw.update(begin=int(c.inp["begin"]))
        extra = list(extra) + ["--hdf5"]
    elif c.fmt == "ufxc":  # a UFXC event file (io/ufxc.cpp), read through --ufxc
        np.asarray(c.inp["words"], "<u4").tofile(imm)
        extra = list(extra) + ["--ufxc"]
    elif c.fmt == "rigaku":  # a Rigaku event file (io/rigaku.cpp), read through --rigaku
        np.asarray(c.inp["words"], "<u8").tofile(imm)
        extra = list(extra) + ["--rigaku"]
    else:
        pkg.synth.write_imm_sparse(imm, h, w, c.inp["off"], c.inp["idx"], c.inp["val"])
    if c.kind == "twotime":
        kw.update(twotime=dict(qbins=[int(q) for q in c.inp["qbins"]], wsize=int(c.inp["wsize"]),
                               method={"staticmap": "StaticMap"}.get(c.method, c.method), filter=str(c.inp["filt"])))
        if "framethreading" in c.name:
            extra = list(extra) + ["--frame_threading"]
    frames_todo = c.F_raw
    if cfg:  # overrides of the configuration (frame range of the job)
        frames_todo = cfg.pop("frames", frames_todo)
        kw.update(cfg)
    cfg = str(tmp_path / "config.hdf5")
    f = pkg.h5lite.File()
    for path, value in refdrv.config_items(c.dq, c.sq, frames_todo, imm, **kw)[0]:
        f.put(path, value)
    f.save(cfg)
    f.close()
    p = subprocess.run([corr, cfg, "--g2out", "--darkout"] + list(extra), stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, env=env)
    assert p.returncode == 0, p.stdout[-2000:]
    g = pkg.h5lite.File(cfg)
    res = g.walk("/exchange")
    assert g.get("/xpcs/delays_per_level")[0, 0] == c.dpl   # the configuration survives the rewrite
    g.close()
    return res, p.stdout


@pytest.mark.parametrize("name", G.names())
def test_corr_matches_reference_result_file(pkg, tmp_path, name):
    c = G.Case(name)
    res, log = _run_corr(pkg, c, tmp_path)
    assert sorted(res) == sorted(c.ref), "result dataset names differ from the reference's"
    for k, ref in c.ref.items():
        got = res[k]
        assert got.shape == ref.shape, "%s: shape %s vs reference %s" % (k, got.shape, ref.shape)
        assert got.dtype == ref.dtype, "%s: dtype %s vs reference %s" % (k, got.dtype, ref.dtype)
        if c.fmt == "rigaku" and k.startswith("timestamp_"):
            # the reference writes [2][raw frames] from the reader's `frames`-long arrays (main.cpp:399-411):
            # only the first `frames` entries of the first row are defined
            got, ref = got[0][: c.F], ref[0][: c.F]
        err, nanmis = G.rel_err(got, ref)
        assert nanmis == 0 and err <= RTOL, "%s: worst relative error %.3g" % (k, err)
    for stage in ("Loading data", "Total"):
        assert stage + " took" in log


def test_corr_integer_fixture_is_bit_exact(pkg, tmp_path):
    c = G.Case("sparse_staletail_32x32")
    res, _ = _run_corr(pkg, c, tmp_path)
    for k in ("G2", "IP", "IF", "norm-0-g2", "pixelSum", "frameSum", "partition-mean-total",
              "partition-mean-partial", "tau", "partition_norm_factor", "timestamp_clock", "timestamp_tick"):
        assert G.n_diff(res[k], c.ref[k]) == 0, k


def test_corr_positional_imm_and_no_compat(pkg, tmp_path):
    """`corr config.hdf5 data.imm` (README form) and --no_compat (exact pair sums)."""
    c = G.Case("staletail_hand_example")
    res, _ = _run_corr(pkg, c, tmp_path, extra=[str(tmp_path / "data.imm"), "--no_compat"])
    k = list(c.ref["tau"].ravel()).index(40.0)
    assert c.ref["G2"][k, 0] == 0.0                                  # the reference drops the pair
    assert res["G2"][k, 0] == np.float32(0.0625) / np.float32(118)   # the exact value


def test_corr_frameout(pkg, tmp_path):
    """--frameout=N (main.cpp:276-310): the first N post-filter frames as dataset frames_out, declared
    (height, width, N) and filled [N][pixels] like the reference's buffer."""
    c = G.Case("sparse_staletail_32x32")
    N = 7
    res, _ = _run_corr(pkg, c, tmp_path, extra=["--frameout=%d" % N])
    h, w = c.dq.shape
    fo = res["frames_out"]
    assert fo.shape == (h, w, N) and fo.dtype == np.float32
    got = fo.ravel().reshape(N, h * w)
    off, idx, val = c.inp["off"], c.inp["idx"], c.inp["val"]
    valid = ((c.dq.ravel() > 0) & (c.sq.ravel() > 0))
    want = np.zeros((N, h * w), np.float32)
    for f in range(N):
        sl = slice(int(off[f]), int(off[f + 1]))
        np.add.at(want[f], idx[sl], val[sl].astype(np.float32))
    want *= valid[None, :]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", ["hdf5_stack_24x32", "hdf5_stack_u32_24x32"])
def test_corr_hdf5_chunked_deflate_shuffle_stack(pkg, tmp_path, name):
    """--hdf5 on a frame stack stored the way detector files are: one chunk per frame, byte shuffle + deflate
    (io/hdf5.cpp:62-221 reads through H5Dread, which decodes them; h5lite's chunk reader does here).  Same results
    as the reference produced from the contiguous stack of the fixture."""
    c = G.Case(name)
    h, w = c.dq.shape
    res, _ = _run_corr(pkg, c, tmp_path, stack_storage=dict(chunk=(1, h, w), deflate=6, shuffle=True))
    for k in ("G2", "IP", "IF", "norm-0-g2", "pixelSum", "frameSum"):
        assert G.n_diff(res[k], c.ref[k]) == 0, k


def test_corr_twotime_matrix_is_stored_as_one_deflate_chunk(pkg, tmp_path):
    """C2T_all/g2_* leave `corr` the way the reference stores them (write2DData(..., compression = true)): chunked +
    deflate; h5lite reads them back (values checked against the fixture by the parametrised test above)."""
    c = G.Case("twotime_symmetric_none")
    res, _ = _run_corr(pkg, c, tmp_path)
    g = pkg.h5lite.File(str(tmp_path / "config.hdf5"))
    notes = g.report("notes")
    g.close()
    assert any("C2T_all/g2_00001" in n and "chunked + filtered" in n for n in notes), notes


def test_corr_outfile_leaves_the_input_untouched(pkg, tmp_path):
    """--outfile=PATH: configuration + results go to PATH, the user's file is not rewritten."""
    c = G.Case("sparse_int_24x24")
    out = str(tmp_path / "results.hdf5")
    corr = os.path.join(os.path.dirname(pkg.cabi.LIB_PATH), "corr")
    imm = str(tmp_path / "data.imm")
    h, w = c.dq.shape
    pkg.synth.write_imm_sparse(imm, h, w, c.inp["off"], c.inp["idx"], c.inp["val"])
    cfg = str(tmp_path / "config.hdf5")
    f = pkg.h5lite.File()
    for path, value in refdrv.config_items(c.dq, c.sq, c.F_raw, imm, dpl=c.dpl, static_window=c.swindow)[0]:
        f.put(path, value)
    f.save(cfg)
    f.close()
    before = open(cfg, "rb").read()
    p = subprocess.run([corr, cfg, "--outfile=" + out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout[-2000:]
    assert open(cfg, "rb").read() == before
    g = pkg.h5lite.File(out)
    res = g.walk("/exchange")
    assert g.get("/xpcs/delays_per_level")[0, 0] == c.dpl
    g.close()
    assert G.n_diff(res["norm-0-g2"], c.ref["norm-0-g2"]) == 0
