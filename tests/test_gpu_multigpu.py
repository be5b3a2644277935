"""Multi-GPU path: one handle per GPU, frame slabs in, pixel shards out, NCCL inside the library (csrc/comm.cu).

* one GPU: the slab partition kernels (k_demux_count / k_demux_scan / k_demux_scatter / k_merge_offsets) with a
  one-rank communicator and XPCS_SLAB_FORCE_EXCHANGE -- everything except the NVLink transfer itself;
* two or more GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`): N threads, N handles, the real
  exchange -- every output must equal the single-GPU run bit for bit (integer counts), also `corr --gpus N`.
"""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import golden_util as G
from conftest import make_case

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refdrv  # noqa: E402  (configuration key list)

pytestmark = pytest.mark.gpu


def _device_count():
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cu.cuInit(0) != 0 or cu.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        return 0


def _single(pkg, dq, sq, F, off, idx, val, **kw):
    c = pkg.Correlator(dq, sq, F, device=0, **kw)
    c.push_sparse(idx, val, off)
    sums = c.finish_ingest()
    Gs = c.multitau()
    g2, se = c.normalize()
    c.close()
    return sums, Gs, g2, se


def _cuts(off, n):
    """frame cuts balanced by event count"""
    E = int(off[-1])
    cut = [0]
    for r in range(1, n):
        f = int(np.searchsorted(off, E * r // n, side="left"))
        cut.append(min(max(f, cut[-1]), len(off) - 1))
    cut.append(len(off) - 1)
    return cut


def _sharded(pkg, n, dq, sq, F, off, idx, val, want_g=True, rounds=1, **kw):
    """n threads = n ranks; returns rank results list of (sums, (G2, IP, IF) or None, g2, se, kernel report, rows,
    transport).  rounds > 1 repeats the job on the same handles (buffers and peer mappings are reused)."""
    uid = pkg.comm_unique_id()
    cut = _cuts(off, n)
    out = [None] * n
    err = [None] * n

    def worker(r):
        try:
            c = pkg.Correlator(dq, sq, F, device=r, shard_index=r, shard_count=n, **kw)
            c.comm_init(n, r, uid)
            a, b = cut[r], cut[r + 1]
            for _ in range(rounds):
                c.reset()
                c.push_sparse_slab(a, idx, val, off[a: b + 1])
                sums = c.finish_ingest()
                Gs = c.multitau() if want_g else c.multitau(want=False)
                g2, se = c.normalize()
            out[r] = (sums, Gs, g2, se, c.kernel_report(), c.info().n_rows, c.comm_transport())
            c.close()
        except Exception as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=worker, args=(r,)) for r in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in th), "a rank hangs in a collective"
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.mark.parametrize("runs", ["0", "1"])
def test_slab_partition_kernels_on_one_gpu(pkg, monkeypatch, runs):
    """A one-rank communicator with the exchange forced: the slab goes through count / scan / scatter / merge and
    must give exactly what a plain push gives -- with the event-by-event scatter and with the run-gathering one
    (the kernel the direct NVLink transport uses from four ranks on; frames of up to 3 000 events here, so that a
    frame spans several 1 024-event chunks)."""
    if pkg.cabi.load().xpcs_comm_nccl_version() == 0:
        pytest.skip("NCCL not loadable")
    monkeypatch.setenv("XPCS_DEMUX_RUNS", runs)
    dq, sq, off, idx, val = make_case(pkg, 96, 80, 300, 0.35, 21, n_dynamic=5, static_per_dynamic=3)
    F = 300
    ref = _single(pkg, dq, sq, F, off, idx, val)
    monkeypatch.setenv("XPCS_SLAB_FORCE_EXCHANGE", "1")
    c = pkg.Correlator(dq, sq, F, device=0)
    c.comm_init(1, 0, pkg.comm_unique_id())
    c.push_sparse_slab(0, idx, val, off)
    sums = c.finish_ingest()
    Gs = c.multitau()
    g2, se = c.normalize()
    rep = c.kernel_report()
    c.close()
    assert rep.get("k_demux_scatter", (0, 0))[1] == 1 and rep.get("k_merge_offsets", (0, 0))[1] == 1
    for k in ("pixel_sum", "frame_sum", "part_total", "part_partial"):
        assert G.n_diff(sums[k], ref[0][k]) == 0, k
    for a, b in zip(Gs, ref[1]):
        assert np.array_equal(a, b)
    assert np.array_equal(g2, ref[2], equal_nan=True) and np.array_equal(se, ref[3], equal_nan=True)


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("n", [2, 4, 8])
def test_sharded_equals_single_gpu(pkg, n):
    if _device_count() < n:
        pytest.skip("needs %d GPUs" % n)
    dq, sq, off, idx, val = make_case(pkg, 96, 80, 3000, 0.01, 33, n_dynamic=6, static_per_dynamic=4)
    F = 3000
    ref = _single(pkg, dq, sq, F, off, idx, val)
    res = _sharded(pkg, n, dq, sq, F, off, idx, val)
    assert sum(r[5] for r in res) == int(((dq > 0) & (sq > 0)).sum())
    Gsum = [np.zeros_like(ref[1][0]) for _ in range(3)]
    for r in range(n):
        sums, Gs, g2, se, rep, _, transport = res[r]
        # whole-detector sums and g2 on EVERY rank, bit-identical to one GPU
        for k in ("pixel_sum", "frame_sum", "part_total", "part_partial"):
            assert G.n_diff(sums[k], ref[0][k]) == 0, (r, k)
        assert np.array_equal(g2, ref[2], equal_nan=True), r
        assert np.array_equal(se, ref[3], equal_nan=True), r
        if transport == 1:   # direct NVLink stores: no staged exchange, one barrier
            assert rep.get("k_demux_scatter_p2p", (0, 0))[1] == 1 and rep.get("nccl_barrier", (0, 0))[1] == 1
            assert "nccl_exchange" not in rep
        else:
            assert rep.get("nccl_exchange", (0, 0))[1] == 1
        assert rep.get("nccl_allreduce", (0, 0))[1] >= 2
        for k in range(3):
            Gsum[k] += Gs[k]   # every pixel has one owner, the others hold zeros
    for k in range(3):
        assert np.array_equal(Gsum[k], ref[1][k])


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_framesum_normalisation_and_flatfield(pkg, oracle):
    """normalize_by_framesum needs the frame sums of ALL pixels before the rows are scaled: all-reduced inside the
    ingest.  Float values: within 1e-5 of the single-GPU run."""
    n = 2
    dq, sq, off, idx, val = make_case(pkg, 64, 64, 1500, 0.02, 35, n_dynamic=4, static_per_dynamic=3)
    F = 1500
    flat = pkg.synth.flatfield(64 * 64)
    kw = dict(flatfield=flat, normalize_by_framesum=True)
    ref = _single(pkg, dq, sq, F, off, idx, val, **kw)
    res = _sharded(pkg, n, dq, sq, F, off, idx, val, **kw)
    for r in range(n):
        sums, Gs, g2, se, rep, _, _ = res[r]
        for k in ("pixel_sum", "frame_sum", "part_total", "part_partial"):
            err, nanmis = G.rel_err(sums[k], ref[0][k])
            assert nanmis == 0 and err <= 1e-5, (r, k, err)
        err, nanmis = G.rel_err(g2, ref[2])
        assert nanmis == 0 and err <= 1e-5, (r, err)


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_corr_gpus_flag_equals_single_gpu_file(pkg, tmp_path):
    """`corr config.hdf5 --imm data.imm --g2out --gpus N`: the host program shards the job itself."""
    n = min(_device_count(), 4)
    dq, sq, off, idx, val = make_case(pkg, 64, 48, 2000, 0.015, 37, n_dynamic=5, static_per_dynamic=3)
    F = 2000
    corr = os.path.join(os.path.dirname(pkg.cabi.LIB_PATH), "corr")
    imm = str(tmp_path / "data.imm")
    pkg.synth.write_imm_sparse(imm, 64, 48, off, idx, val)
    results = []
    for gpus in (1, n):
        cfg = str(tmp_path / ("config%d.hdf5" % gpus))
        f = pkg.h5lite.File()
        for path, value in refdrv.config_items(dq, sq, F, imm, dpl=8)[0]:
            f.put(path, value)
        f.save(cfg)
        f.close()
        p = subprocess.run([corr, cfg, "--g2out", "--gpus", str(gpus)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-3000:]
        g = pkg.h5lite.File(cfg)
        results.append(g.walk("/exchange"))
        g.close()
    one, many = results
    assert sorted(one) == sorted(many)
    for k in one:
        assert one[k].shape == many[k].shape and one[k].dtype == many[k].dtype, k
        assert G.n_diff(one[k], many[k]) == 0, k


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_ragged_slabs(pkg):
    """Edge cases of the exchange: a rank whose slab has no frames at all, a rank whose frames hold no events, frames
    without events in the middle, and a job with fewer frames than would fill every rank."""
    n = 2
    dq, sq, off, idx, val = make_case(pkg, 32, 32, 40, 0.02, 39, n_dynamic=3, static_per_dynamic=2)
    F = 40
    # frames 10..19 empty
    keep = np.ones(idx.size, bool)
    keep[int(off[10]):int(off[20])] = False
    cnt = np.diff(off)
    cnt[10:20] = 0
    idx, val = idx[keep], val[keep]
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    ref = _single(pkg, dq, sq, F, off, idx, val)
    uid = pkg.comm_unique_id()
    out, err = [None] * n, [None] * n
    cuts = {"empty_rank": [0, 0, F], "empty_frames_rank": [0, 10, F]}
    for name, cut in cuts.items():
        uid = pkg.comm_unique_id()

        def worker(r):
            try:
                c = pkg.Correlator(dq, sq, F, device=r, shard_index=r, shard_count=n)
                c.comm_init(n, r, uid)
                a, b = cut[r], cut[r + 1]
                c.push_sparse_slab(a, idx, val, off[a: b + 1])
                sums = c.finish_ingest()
                c.multitau(want=False)
                out[r] = (sums, c.normalize())
                c.close()
            except Exception as e:  # noqa: BLE001
                err[r] = e

        th = [threading.Thread(target=worker, args=(r,)) for r in range(n)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=120)
        assert not any(t.is_alive() for t in th), name
        for e in err:
            if e is not None:
                raise e
        for r in range(n):
            for k in ("pixel_sum", "frame_sum", "part_total", "part_partial"):
                assert G.n_diff(out[r][0][k], ref[0][k]) == 0, (name, r, k)
            assert np.array_equal(out[r][1][0], ref[2], equal_nan=True), (name, r)


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("staged", [False, True])
def test_both_transports_and_buffer_reuse(pkg, monkeypatch, staged):
    """The slab exchange over direct NVLink stores (default inside one process: peer access) and over the staged
    ncclSend/ncclRecv path (XPCS_NO_P2P), three jobs in a row on the same handles, the second one larger than the
    first (the receive buffers grow and the peers' mappings must follow): always the single-GPU result."""
    if staged:
        monkeypatch.setenv("XPCS_NO_P2P", "1")
    n = 2
    dq, sq, off, idx, val = make_case(pkg, 64, 64, 2500, 0.012, 41, n_dynamic=4, static_per_dynamic=3)
    F = 2500
    ref = _single(pkg, dq, sq, F, off, idx, val)
    # a smaller job first (half the events dropped), then the full one twice
    keep = np.arange(idx.size) % 2 == 0
    cnt = np.add.reduceat(keep.astype(np.int64), off[:-1].clip(max=idx.size - 1)) * (np.diff(off) > 0)
    off_small = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    uid = pkg.comm_unique_id()
    out, err = [None] * n, [None] * n
    cut = _cuts(off, n)

    def worker(r):
        try:
            c = pkg.Correlator(dq, sq, F, device=r, shard_index=r, shard_count=n)
            c.comm_init(n, r, uid)
            a, b = cut[r], cut[r + 1]
            c.push_sparse_slab(a, idx[keep], val[keep], off_small[a: b + 1])
            c.finish_ingest(want=False)
            c.multitau(want=False)
            c.normalize()
            for _ in range(2):
                c.reset()
                c.push_sparse_slab(a, idx, val, off[a: b + 1])
                sums = c.finish_ingest()
                c.multitau(want=False)
                g2, se = c.normalize()
            out[r] = (sums, g2, se, c.comm_transport(), c.kernel_report())
            c.close()
        except Exception as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=worker, args=(r,)) for r in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in th)
    for e in err:
        if e is not None:
            raise e
    for r in range(n):
        sums, g2, se, transport, rep = out[r]
        assert transport == (0 if staged else 1), "transport %d" % transport
        for k in ("pixel_sum", "frame_sum", "part_total", "part_partial"):
            assert G.n_diff(sums[k], ref[0][k]) == 0, (r, k)
        assert np.array_equal(g2, ref[2], equal_nan=True) and np.array_equal(se, ref[3], equal_nan=True)


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_corr_twotime_partitions_over_gpus(pkg, tmp_path):
    """Two-time with --gpus 2: the listed dynamic partitions are independent (corr.cpp:799) and are spread over the
    GPUs, one handle per GPU; the result file equals the one-GPU file and the reference's fixture."""
    c = G.Case("twotime_staticmap_none")
    corr = os.path.join(os.path.dirname(pkg.cabi.LIB_PATH), "corr")
    imm = str(tmp_path / "data.imm")
    h, w = c.dq.shape
    pkg.synth.write_imm_sparse(imm, h, w, c.inp["off"], c.inp["idx"], c.inp["val"])
    results = []
    for gpus in (1, 2):
        cfg = str(tmp_path / ("tt%d.hdf5" % gpus))
        f = pkg.h5lite.File()
        kw = dict(dpl=c.dpl, static_window=c.swindow,
                  twotime=dict(qbins=[int(q) for q in c.inp["qbins"]], wsize=int(c.inp["wsize"]), method="StaticMap",
                               filter=str(c.inp["filt"])))
        for path, value in refdrv.config_items(c.dq, c.sq, c.F_raw, imm, **kw)[0]:
            f.put(path, value)
        f.save(cfg)
        f.close()
        p = subprocess.run([corr, cfg, "--gpus", str(gpus)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-3000:]
        if gpus > 1:
            assert "dynamic partitions over 2 GPUs" in p.stdout
        g = pkg.h5lite.File(cfg)
        results.append(g.walk("/exchange"))
        g.close()
    one, two = results
    assert sorted(one) == sorted(two) == sorted(c.ref)
    for k in one:
        assert G.n_diff(one[k], two[k]) == 0, k
        err, nanmis = G.rel_err(two[k], c.ref[k])
        assert nanmis == 0 and err <= 1e-5, (k, err)
