"""The HOST logic of the drop-in program `corr`, on a machine without a GPU.

`corr` (xpcs-eigen_b200/host/corr_main.cpp) is linked against libxpcs_b200.so with run path $ORIGIN.  Here a copy of
the binary sits next to a RECORDER (tests/host_mt/cabi_recorder.cpp) that implements the same C-ABI symbols, computes
nothing, writes down what `corr` hands over and answers result requests with zeros of the right shape.  Checked:
configuration parsing and the readers (the events that reach the boundary are exactly the events of the input, frame
by frame, with their timestamps), the call sequence of the resident and of the streamed job, the cut of a stream into
pushes, and the names / shapes / types of the result datasets (reference main.cpp:345-426, h5_result.cpp:56-347).
No number `corr` writes in such a run is compared with anything -- parity is the business of the -m gpu tests."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import golden_util as G
from test_gpu_corr_host import _run_corr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def recorder_dir(pkg, tmp_path_factory):
    d = tmp_path_factory.mktemp("corr_recorder")
    corr = os.path.join(os.path.dirname(pkg.cabi.LIB_PATH), "corr")
    if not os.path.exists(corr):
        pytest.skip("corr is not built (python -c 'import __graft_entry__ as g; g.build()')")
    shutil.copy2(corr, str(d / "corr"))
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(d / "libxpcs_b200.so"),
                           os.path.join(ROOT, "tests", "host_mt", "cabi_recorder.cpp")])
    return d


def run(pkg, recorder_dir, case, tmp_path, extra=(), cfg=None):
    rec = tmp_path / "record"
    rec.mkdir()
    env = dict(os.environ, XPCS_RECORD_DIR=str(rec))
    res, log = _run_corr(pkg, case, tmp_path, extra=extra, corr_path=str(recorder_dir / "corr"), env=env, cfg=cfg)
    calls = open(str(rec / "calls.txt")).read().splitlines()
    got = dict(idx=np.fromfile(str(rec / "idx.bin"), np.int32), val=np.fromfile(str(rec / "val.bin"), np.int16),
               frame_events=np.fromfile(str(rec / "frame_events.bin"), np.int64),
               clock=np.fromfile(str(rec / "clock.bin"), np.float64), ticks=np.fromfile(str(rec / "ticks.bin"), np.float64))
    return res, log, calls, got


def check_events(case, got):
    off, idx, val = case.inp["off"], case.inp["idx"], case.inp["val"]
    F = case.F_raw
    assert np.array_equal(got["frame_events"], np.diff(off[: F + 1])), "events per frame"
    assert np.array_equal(got["idx"], idx[: off[F]]) and np.array_equal(got["val"], val[: off[F]]), "payload"
    assert got["clock"].size == F and got["ticks"].size == F


def check_result_layout(case, res):
    """dataset names of the reference's result file; shapes and types as the reference wrote them"""
    assert sorted(res) == sorted(case.ref), (sorted(res), sorted(case.ref))
    for k, ref in case.ref.items():
        assert res[k].shape == ref.shape and res[k].dtype == ref.dtype, k
    assert np.array_equal(res["tau"], case.ref["tau"])          # the delay list is host work (xpcs_delay_schedule)
    assert np.array_equal(res["timestamp_clock"][0], case.ref["timestamp_clock"][0])


def test_the_recorder_runs_the_reference_schedule(pkg, recorder_dir):
    import ctypes as C
    rec = C.CDLL(str(recorder_dir / "libxpcs_b200.so"))
    lib = pkg.cabi.load()
    for F, dpl in ((15, 8), (16, 8), (17, 8), (33, 8), (600, 8), (1000, 8), (9999, 8), (10000, 4), (100000, 8), (1000000, 8)):
        a = [np.zeros(512, np.int32) for _ in range(4)]
        n1 = rec.xpcs_delay_schedule(F, dpl, a[0].ctypes.data_as(C.c_void_p), a[1].ctypes.data_as(C.c_void_p), 512)
        n2 = lib.xpcs_delay_schedule(F, dpl, a[2].ctypes.data, a[3].ctypes.data, 512)
        assert n1 == n2 and np.array_equal(a[0], a[2]) and np.array_equal(a[1], a[3]), (F, dpl)


@pytest.mark.parametrize("name", ["sparse_staletail_32x32", "sparse_odd_dpl4", "sparse_int_24x24"])
def test_resident_job(pkg, recorder_dir, tmp_path, name):
    case = G.Case(name)
    res, log, calls, got = run(pkg, recorder_dir, case, tmp_path)
    check_events(case, got)
    check_result_layout(case, res)
    names = [c.split()[0] for c in calls]
    assert names == ["create", "push_sparse", "finish_ingest", "multitau", "normalize"], calls
    assert "compat_flags=1" in calls[0]                      # the reference's behaviour is the default
    assert calls[1] == "push_sparse nframes=%d events=%d" % (case.F_raw, got["idx"].size)
    assert "g2out=1" in calls[3]
    for stage in ("Loading data", "Total"):
        assert stage + " took" in log


@pytest.mark.parametrize("name,K", [("sparse_staletail_32x32", 64), ("sparse_staletail_32x32", 1024), ("sparse_odd_dpl4", 128),
                                    ("sparse_int_24x24", 8192)])
def test_streamed_job(pkg, recorder_dir, tmp_path, name, K):
    """--stream_frames K: the stale-tail flag is cleared (and said so), the stream is opened with K, the frames arrive in
    order in pushes that start on chunk boundaries, nothing is read twice or dropped, the outputs are requested as for
    a resident job"""
    case = G.Case(name)
    res, log, calls, got = run(pkg, recorder_dir, case, tmp_path, extra=["--stream_frames", str(K)])
    check_events(case, got)
    check_result_layout(case, res)
    names = [c.split()[0] for c in calls]
    assert names[0] == "create" and "compat_flags=0" in calls[0]
    assert calls[1] == "stream_begin chunk_frames=%d" % K
    pushes = [c for c in calls if c.startswith("stream_push_sparse")]
    assert names == ["create", "stream_begin"] + ["stream_push_sparse"] * len(pushes) + ["stream_finish", "multitau", "normalize"]
    nfr = [int(c.split()[1].split("=")[1]) for c in pushes]
    assert sum(nfr) == case.F_raw
    assert all(n % K == 0 for n in nfr[:-1]), "only the last push may end inside a chunk"
    # a compressed IMM file is read chunk by chunk: the host never holds more than one chunk of frames
    assert all(n <= K for n in nfr) and len(pushes) == (case.F_raw + K - 1) // K
    assert all(c.endswith("first_offset=0") for c in pushes)
    assert "exact multi-tau sums" in log


@pytest.mark.parametrize("name,flag", [("rigaku_compact_32x40", "--rigaku"), ("ufxc_wrap_48x40", "--ufxc")])
def test_streamed_job_from_event_word_files(pkg, recorder_dir, tmp_path, name, flag):
    """Rigaku / UFXC files are decoded as a whole (event words in file order); the stream then takes one push that the
    library cuts into chunks.  The events that arrive are those of the resident run of the same file."""
    case = G.Case(name)
    a, b = tmp_path / "stream", tmp_path / "resident"
    a.mkdir()
    b.mkdir()
    res, log, calls, got = run(pkg, recorder_dir, case, a, extra=["--stream_frames", "64"])
    _, _, calls_r, got_r = run(pkg, recorder_dir, case, b)
    for k in got:
        assert np.array_equal(got[k], got_r[k]), k
    pushes = [c for c in calls if c.startswith("stream_push_sparse")]
    assert len(pushes) == 1 and "nframes=%d " % case.F_raw in pushes[0]
    assert [c.split()[0] for c in calls_r] == ["create", "push_sparse", "finish_ingest", "multitau", "normalize"]
    assert sorted(res) == sorted(case.ref)


def test_stream_refused_for_jobs_it_cannot_run(pkg, recorder_dir, tmp_path):
    """two-time jobs, --frameout and --gpus N keep the resident path: corr says so and stops before touching the library"""
    case = G.Case("sparse_staletail_32x32")
    with pytest.raises(AssertionError) as e:
        run(pkg, recorder_dir, case, tmp_path, extra=["--stream_frames", "64", "--frameout", "3"])
    assert "--stream_frames is for multi-tau jobs" in str(e.value)


@pytest.mark.parametrize("name", G.names())
def test_every_fixture_job_has_the_reference_result_layout(pkg, recorder_dir, tmp_path, name):
    """every job kind the golden fixtures cover -- sparse / dense IMM, UFXC, Rigaku (stride, average), HDF5 stacks, two-time
    with symmetric and StaticMap smoothing -- through the host program: result dataset names, shapes and types are the
    ones the reference binary wrote (tests/golden), and the library is called in the order the C-ABI prescribes"""
    case = G.Case(name)
    res, log, calls, got = run(pkg, recorder_dir, case, tmp_path)
    assert sorted(res) == sorted(case.ref), (sorted(set(res) ^ set(case.ref)))
    for k, ref in case.ref.items():
        assert res[k].shape == ref.shape, "%s: shape %s vs reference %s" % (k, res[k].shape, ref.shape)
        assert res[k].dtype == ref.dtype, "%s: dtype %s vs reference %s" % (k, res[k].dtype, ref.dtype)
    names = [c.split()[0] for c in calls]
    assert names[0] == "create"
    if case.kind == "dense":
        h, w = case.dq.shape
        frames = np.asarray(case.inp["frames"], np.int16).reshape(-1, h * w)
        dark = np.fromfile(str(tmp_path / "record" / "dark.bin"), np.int16)
        dense = np.fromfile(str(tmp_path / "record" / "dense.bin"), np.int16)
        nd = case.darks or 0
        assert ("set_dark" in names) == (nd > 0)
        assert np.array_equal(dark, frames[:nd].ravel())                       # main.cpp:227-239: darks from the file start
        assert np.array_equal(dense, frames[nd: nd + case.F_raw].ravel())
        assert names[-3:] == ["finish_ingest", "multitau", "normalize"] or "get_dark" in names
    elif case.kind == "twotime":
        assert "finish_ingest" in names and names.count("twotime") == len(case.inp["qbins"])
        assert "multitau" not in names
    else:
        assert names[1] == "push_sparse" and names[2] == "finish_ingest"
        assert got["frame_events"].size == case.F_raw


def test_sharded_job(pkg, recorder_dir, tmp_path):
    """corr --gpus 3: one handle per shard, every shard gets a slab of consecutive frames of the whole detector, the slabs
    tile the frame range, all ranks join one communicator and make the same calls"""
    case = G.Case("sparse_staletail_32x32")
    res, log, calls0, _ = None, None, None, None
    rec = tmp_path / "record"
    rec.mkdir()
    env = dict(os.environ, XPCS_RECORD_DIR=str(rec))
    res, log = _run_corr(pkg, case, tmp_path, extra=["--gpus", "3"], corr_path=str(recorder_dir / "corr"), env=env)
    assert sorted(res) == sorted(case.ref)
    off = case.inp["off"]
    nxt, total = 0, 0
    for r in range(3):
        d = rec / ("shard%d" % r)
        calls = open(str(d / "calls.txt")).read().splitlines()
        names = [c.split()[0] for c in calls]
        assert names == ["create", "comm_init", "push_sparse_slab", "finish_ingest", "multitau", "normalize"], calls
        assert "nranks=3 rank=%d id0=90" % r in calls[1]
        f = dict(kv.split("=") for kv in calls[2].split()[1:])
        assert int(f["first"]) == nxt, "slabs must tile the frame range in rank order"
        fe = np.fromfile(str(d / "frame_events.bin"), np.int64)
        assert fe.size == int(f["nframes"]) and np.array_equal(fe, np.diff(off[nxt: nxt + fe.size + 1]))
        idx = np.fromfile(str(d / "idx.bin"), np.int32)
        assert np.array_equal(idx, case.inp["idx"][off[nxt]: off[nxt + fe.size]])
        nxt += fe.size
        total += idx.size
    assert nxt == case.F_raw and total == off[case.F_raw]


@pytest.mark.parametrize("extra", [(), ("--stream_frames", "64")])
def test_frame_range_of_the_job(pkg, recorder_dir, tmp_path, extra):
    """data_begin_todo / data_end_todo select frames 301 .. 1500 of the file (main.cpp:241-245 skips the first 300): exactly
    those frames reach the boundary, whole or chunk by chunk"""
    case = G.Case("sparse_staletail_32x32")
    res, log, calls, got = run(pkg, recorder_dir, case, tmp_path, extra=list(extra), cfg=dict(begin=301, frames=1200))
    off, idx, val = case.inp["off"], case.inp["idx"], case.inp["val"]
    assert np.array_equal(got["frame_events"], np.diff(off[300:1501]))
    assert np.array_equal(got["idx"], idx[off[300]: off[1500]]) and np.array_equal(got["val"], val[off[300]: off[1500]])
    assert "frames=1200" in calls[0]
    assert res["frameSum"].shape == (2, 1200) and res["G2"].shape[0] == res["tau"].shape[1]
