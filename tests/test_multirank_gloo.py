"""N>1 host-side logic on CPU: two processes over gloo (127.0.0.1).
Covers what bench.py / a multi-GPU `corr` does around the kernels: every rank plans its pixel
shard through the host-only C-ABI call xpcs_plan_shard (static-partition aligned, disjoint,
covering, balanced), lays its normalisation partials out in the global [3][nseg][T] buffer
with zeros for foreign segments, and one all_reduce(SUM) yields the single-shard buffer bit for
bit on every rank.  The device-side counterpart (3 shards on one GPU) is in
test_gpu_parity.py::test_shards_sum_to_single_gpu_partials."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _segment_values(nseg, T):
    """Deterministic stand-in for the per-segment partial rows a shard's kernels would write."""
    s = np.arange(nseg, dtype=np.float64)[:, None]
    t = np.arange(T, dtype=np.float64)[None, :]
    return np.stack([1.0 + 0.001 * s + 1e-5 * t, 10.0 * s + t, 100.0 * s + t * t])  # [3][nseg][T]


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        import __graft_entry__ as entry
        pkg = entry.load_package()
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dq, sq = pkg.synth.annular_qmaps(96, 80, n_dynamic=5, static_per_dynamic=4, r_min=3.0)
        F = 700
        full, rows_all = pkg.plan_shard(dq, sq, F, 0, 1)
        plan, rows = pkg.plan_shard(dq, sq, F, rank, world)
        T, nseg = plan.n_delays, plan.n_segments
        assert (nseg, T, plan.n_rows_total) == (full.n_segments, full.n_delays, full.n_rows_total)
        # (1) shards are disjoint and cover the single-shard row list in order
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([rows.size], dtype=torch.int64))
        sizes = [int(x) for x in sizes]
        assert sum(sizes) == rows_all.size
        start = sum(sizes[:rank])
        assert np.array_equal(rows, rows_all[start: start + rows.size])
        # static-partition aligned: no static bin is split between ranks
        mine = set(np.unique(sq.ravel()[rows]).tolist())
        bins = [None] * world
        dist.all_gather_object(bins, mine)
        for r in range(world):
            if r != rank:
                assert not (mine & bins[r]), "static partition split across ranks"
        # balanced within one segment's worth of pixels
        assert abs(rows.size - rows_all.size / world) <= rows_all.size / nseg * 2 + 1
        # segment ranges tile [0, nseg)
        cuts = [None] * world
        dist.all_gather_object(cuts, (plan.seg_first, plan.seg_last))
        assert cuts[0][0] == 0 and cuts[-1][1] == nseg
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
        # (2) the exchange: foreign entries are zero, SUM over ranks == the single-shard buffer
        ref = _segment_values(nseg, T)
        buf = np.zeros_like(ref)
        buf[:, plan.seg_first: plan.seg_last, :] = ref[:, plan.seg_first: plan.seg_last, :]
        t = torch.from_numpy(buf.reshape(-1).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        assert np.array_equal(t.numpy(), ref.reshape(-1)), "all-reduced partials differ from the single-shard buffer"
        # (3) max-over-ranks timing reduction used by bench.py
        tm = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        assert float(tm) == 10.0 + world - 1
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %r\n%s" % (e, traceback.format_exc())))


@pytest.mark.timeout(180)
def test_two_rank_shard_plan_and_partials_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in sorted(res):
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def test_plan_shard_matches_oracle_qmap(pkg, oracle):
    """The product's host-side BuildQMap equivalent against the oracle restatement of
    configuration.cpp:244-381, incl. a static bin that straddles two dynamic bins."""
    dq, sq = pkg.synth.annular_qmaps(40, 48, n_dynamic=4, static_per_dynamic=3, r_min=2.0)
    sq2 = sq.copy()
    sq2[(dq == 2) & (sq == sq[dq == 2].max())] = sq[dq == 3].min()  # static bin shared by dq 2 and 3
    for d, s in ((dq, sq), (dq, sq2)):
        qm = oracle.QMap(d, s)
        plan, rows = pkg.plan_shard(d, s, 500)
        assert (plan.n_static, plan.n_dynamic, plan.n_segments) == (qm.S, qm.Q, qm.nseg)
        n_mapped = qm.seg_pixels.size
        assert np.array_equal(rows[:n_mapped], qm.seg_pixels)      # (dq, sq, pixel) order
        assert plan.n_rows_total == int((qm.mask != 0).sum())       # orphans still correlated
    # 3 shards tile the row list
    parts = [pkg.plan_shard(dq, sq, 500, k, 3)[1] for k in range(3)]
    assert np.array_equal(np.concatenate(parts), pkg.plan_shard(dq, sq, 500)[1])


def test_reference_arm_only_rank0_prints(tmp_path):
    """bench.py --impl reference under a 2-rank launch: rank 1 exits 0 without output."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
