"""Rank 0 of an N-rank weak-scaling bench run, alone on one GPU (no NCCL): per-kernel times of the shard's job.
usage: python profiles/diag_shard.py [world]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = dict(bench.WORKLOADS["c3"])
pkg = entry.load_package()
dev = torch.device("cuda", 0)
F, occ = wl["F"], wl["occ"]
dq, sq = bench.module_maps(pkg, wl, world)
E_est = int(wl["h"] * wl["w"] * F * occ * 1.02) + 4096
c = pkg.Correlator(dq, sq, F, dpl=8, compat=True, device=0, shard_index=0, shard_count=world, reserve_events=E_est)
stream = torch.cuda.Stream(device=dev)
c.set_stream(stream.cuda_stream)
if world == 1:
    pixels_dev, n_pix = None, dq.size
else:
    own = c.row_pixels()
    masked = np.nonzero((dq.ravel() < 1) | (sq.ravel() < 1))[0].astype(np.int32)[0::world]
    uni = np.sort(np.concatenate([own, masked])).astype(np.int32)
    pixels_dev, n_pix = torch.from_numpy(uni).to(dev), int(uni.size)
d_idx, d_val, d_off = bench.gen_sparse_device(torch, pixels_dev, n_pix, F, occ, 1234, dev)
E = int(d_idx.numel())
print("world", world, "rows", c.info().n_rows, "events", E, "segments local", "?")
def step():
    c.reset()
    c.push_sparse_device(d_idx.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), E, F)
    c.finish_ingest(want=False)
    c.multitau(want=False)
    c.normalize_partials()
    return c.normalize_finish()
for _ in range(3):
    step()
c.kernel_report(reset=True); c.kernel_timing(True)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(5):
    step()
e1.record(stream); torch.cuda.synchronize()
print("ms/step %.3f" % (e0.elapsed_time(e1) / 5))
for k, (ms, n) in sorted(c.kernel_report().items(), key=lambda kv: -kv[1][0])[:7]:
    print("  %-20s %.3f ms/step" % (k, ms / 5))
