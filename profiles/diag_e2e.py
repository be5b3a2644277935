"""Wall-clock split of the host-buffer (e2e) path of the C3 workload: push / finish_ingest / multitau /
normalize, with and without the pipelined ingest.  usage: python profiles/diag_e2e.py [c3|c5] [frames]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
import bench

WL = sys.argv[1] if len(sys.argv) > 1 else "c3"
wl = dict(bench.WORKLOADS[WL])
F = int(sys.argv[2]) if len(sys.argv) > 2 else wl["F"]
wl["F"] = F
pkg = entry.load_package()
dev = torch.device("cuda", 0)
dq, sq = bench.module_maps(pkg, wl, 1)
occ = wl["occ"]
E_est = int(wl["h"] * wl["w"] * F * occ * 1.02) + 4096
d_idx, d_val, d_off = bench.gen_sparse_device(torch, None, dq.size, F, occ, 1234, dev)
E = int(d_idx.numel())
h_idx = torch.empty(E, dtype=torch.int32, pin_memory=True).copy_(d_idx)
h_val = torch.empty(E, dtype=torch.int16, pin_memory=True).copy_(d_val)
h_off = torch.empty(F + 1, dtype=torch.int64, pin_memory=True).copy_(d_off)
torch.cuda.synchronize()
# raw PCIe rate
t0 = time.perf_counter(); d_idx.copy_(h_idx, non_blocking=True); d_val.copy_(h_val, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter() - t0
print("plain H2D of %d MB: %.2f ms (%.1f GB/s)" % (6 * E >> 20, 1e3 * t, 6 * E / t / 1e9))
for mode in ("pipe", "nopipe", "pipe", "nopipe"):
    os.environ.pop("XPCS_NO_PIPELINE", None)
    if mode == "nopipe":
        os.environ["XPCS_NO_PIPELINE"] = "1"
    c = pkg.Correlator(dq, sq, F, dpl=8, compat=True, device=0, reserve_events=E_est)
    stream = torch.cuda.Stream(device=dev)
    c.set_stream(stream.cuda_stream)
    for it in range(4):
        c.reset()
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        c.push_sparse_raw(h_idx.data_ptr(), h_val.data_ptr(), h_off.data_ptr(), F); t.append(time.perf_counter())
        sums = c.finish_ingest(want=True); t.append(time.perf_counter())
        c.multitau(want=False); t.append(time.perf_counter())
        c.normalize_partials(); g2, se = c.normalize_finish(); t.append(time.perf_counter())
        torch.cuda.synchronize(); t.append(time.perf_counter())
        if it >= 2:
            d = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
            print("%-7s push %.2f  finish %.2f  multitau %.2f  normalize %.2f  sync %.2f  total %.2f ms" % (mode, *d, 1e3 * (t[-1] - t[0])))
    rep = c.kernel_report()
    c.close()
