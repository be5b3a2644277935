"""C5 at 5 % occupancy (BASELINE configs[4]'s high end: 2048 x 2048, 5 % -- 2.1e11 events at 1M frames, which no set of
GPUs of one box holds) through the ONLINE multi-tau: the events of a chunk of frames exist on the device only while
that chunk is consumed.

    python profiles/stream_big.py [--h 2048 --w 2048 --frames 65536 --occ 0.05 --chunk 2048] [--out FILE]

Every chunk is generated on the device (bench.py's generator, seed = base + chunk number), pushed with
xpcs_stream_push_sparse_device and dropped.  Timed: the pushes, the finish and the normalisation (CUDA-synchronised
wall clock around each call; the generation is outside).  Checked: G2 / IP / IF of sampled pixel rows bit for bit
against oracle.multitau (exact sums) on the events of those pixels, regenerated chunk by chunk, and norm-0-g2 of
three dynamic bins against oracle.normalize (bench.parity_block).  The cost of a chunk does not depend on how many
chunks came before it, so frames/s of a run with fewer frames than 1M is the rate of the full job."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=2048)
    ap.add_argument("--w", type=int, default=2048)
    ap.add_argument("--frames", type=int, default=65536)
    ap.add_argument("--occ", type=float, default=0.05)
    ap.add_argument("--chunk", type=int, default=2048)
    ap.add_argument("--parity-rows", type=int, default=48)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    pkg = entry.load_package()
    O = entry.load_oracle()
    F, K = a.frames, a.chunk
    dq, sq = pkg.synth.annular_qmaps(a.h, a.w, n_dynamic=36, static_per_dynamic=10, r_min=8.0)
    c = pkg.Correlator(dq, sq, F, dpl=8, compat=False, device=0)
    nchunks = (F + K - 1) // K

    def gen(k):
        nf = min(K, F - k * K)
        return bench.gen_sparse_device(torch, None, a.h * a.w, nf, a.occ, 9000 + k, dev) + (nf,)

    def ev_source():
        for k in range(nchunks):
            i, v, o, nf = gen(k)
            yield i, v, o, k * K

    out = {"what": "stream_big", "h": a.h, "w": a.w, "frames": F, "occupancy": a.occ, "chunk_frames": K, "chunks": nchunks,
           "rows": int(c.info().n_rows), "delays": int(c.T)}
    c.kernel_timing(True)
    c.kernel_report(reset=True)
    t_push = 0.0
    E = 0
    c.stream_begin(K)
    for k in range(nchunks):
        i, v, o, nf = gen(k)
        E += int(i.numel())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c.stream_push_sparse_device(i.data_ptr(), v.data_ptr(), o.data_ptr(), int(i.numel()), nf)
        torch.cuda.synchronize()
        t_push += time.perf_counter() - t0
        del i, v, o
    t0 = time.perf_counter()
    c.stream_finish(want=False)
    c.multitau(want=False)
    g2, se = c.normalize()
    torch.cuda.synchronize()
    t_fin = time.perf_counter() - t0
    rep = c.kernel_report(reset=True)
    c.kernel_timing(False)
    out["events"] = E
    out["event_bytes_frame_major"] = E * 6
    out["push_ms"] = t_push * 1e3
    out["finish_normalize_ms"] = t_fin * 1e3
    out["frames_per_s"] = F / (t_push + t_fin)
    out["events_per_s"] = E / (t_push + t_fin)
    out["ms_per_chunk"] = (t_push * 1e3) / nchunks
    out["kernels_ms"] = {k: round(v[0], 3) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0]) if v[1] > 0}
    out["device_memory_used_gb"] = (torch.cuda.mem_get_info(0)[1] - torch.cuda.mem_get_info(0)[0]) / 1e9
    out["g2_finite"] = bool(np.isfinite(g2).all())
    out["parity"] = bench.parity_block(torch, pkg, O, c, dq, sq, F, ev_source, c.row_pixels(), np.array(g2, copy=True),
                                       n_rows=a.parity_rows, compat=False)
    c.close()
    line = json.dumps(out)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")
    sys.exit(0 if out["parity"]["ok"] else 1)


if __name__ == "__main__":
    main()
