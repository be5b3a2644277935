"""Online multi-tau (xpcs_stream_*) on the GPU: one streamed job against the resident path on the same events.

    python profiles/stream_check.py [--h 1024 --w 1024 --frames 100000 --occ 0.001 --chunk 2048] [--out FILE]

Synthetic events are generated on the device (bench.py's generator), the resident path runs once without the
stale-tail flag (xpcs_push_sparse_device -> finish_ingest -> multitau -> normalize), then the same events are
streamed chunk by chunk from device buffers.  Checked: norm-0-g2 / stderr and the G2 / IP / IF columns of sampled
pixels are bit-identical.  Reported: wall time of both (CUDA-synchronised), the per-kernel CUDA-event times of the
streamed run, the stream kernel's share.  One JSON line on stdout (and in --out)."""
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
    ap.add_argument("--h", type=int, default=1024)
    ap.add_argument("--w", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=100000)
    ap.add_argument("--occ", type=float, default=0.001)
    ap.add_argument("--chunk", type=int, default=2048)
    ap.add_argument("--chunks", default="", help="comma-separated chunk lengths: one JSON line each (resident path run once)")
    ap.add_argument("--profile", action="store_true", help="one streamed pass and nothing else (under ncu)")
    ap.add_argument("--also-bins-only", action="store_true", help="repeat the last chunk length with XPCS_ST_EVENTS=0 (every level walked by its bins)")
    ap.add_argument("--host", action="store_true", help="also time the stream fed from page-locked HOST buffers (xpcs_stream_push_sparse)")
    ap.add_argument("--out", default="")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    pkg = entry.load_package()
    F = a.frames
    dq, sq = pkg.synth.annular_qmaps(a.h, a.w, n_dynamic=36, static_per_dynamic=10, r_min=8.0)
    d_idx, d_val, d_off = bench.gen_sparse_device(torch, None, a.h * a.w, F, a.occ, 4321, dev)
    E = int(d_idx.numel())
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    c = pkg.Correlator(dq, sq, F, dpl=8, compat=False, device=0, reserve_events=E + 1024)
    rng = np.random.default_rng(7)
    sample = np.sort(rng.choice(c.row_pixels(), size=min(4000, c.info().n_rows), replace=False)).astype(np.int32)

    def resident():
        c.reset()
        c.push_sparse_device(d_idx.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), E, F)
        c.finish_ingest(want=False)
        c.multitau(want=False)
        return c.normalize()

    off_h = d_off.cpu().numpy()
    same = lambda x, y: bool(np.array_equal(np.asarray(x).view(np.uint32), np.asarray(y).view(np.uint32)))
    lines, ok = [], True
    g2_r = None
    todo = [(int(x), None) for x in a.chunks.split(",")] if a.chunks else [(a.chunk, None)]
    if a.also_bins_only:
        todo.append((todo[-1][0], "0"))
    for chunk, events_knob in todo:
        os.environ.pop("XPCS_ST_EVENTS", None)
        if events_knob is not None:
            os.environ["XPCS_ST_EVENTS"] = events_knob
        # chunk views: events of frames [f0, f1) and their offsets rebased to 0
        chunks = []
        for f0 in range(0, F, chunk):
            f1 = min(F, f0 + chunk)
            e0, e1 = int(off_h[f0]), int(off_h[f1])
            chunks.append((d_idx[e0:e1], d_val[e0:e1], (d_off[f0:f1 + 1] - e0).contiguous(), e1 - e0, f1 - f0))
        torch.cuda.synchronize()

        def streamed():
            c.reset()
            c.stream_begin(chunk)
            for i, v, o, ne, nf in chunks:
                c.stream_push_sparse_device(i.data_ptr(), v.data_ptr(), o.data_ptr(), ne, nf)
            c.stream_finish(want=False)
            c.multitau(want=False)
            return c.normalize()

        if a.profile:
            streamed()
            torch.cuda.synchronize()
            continue
        out = {"what": "stream_check", "tag": a.tag, "h": a.h, "w": a.w, "frames": F, "occupancy": a.occ, "chunk_frames": chunk,
               "XPCS_ST_EVENTS": events_knob, "events": E, "rows": int(c.info().n_rows), "delays": int(c.T), "chunks": len(chunks)}
        if g2_r is None:
            g2_r, se_r = [np.array(x, copy=True) for x in resident()]
            cor_r = c.correlators(sample)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            resident()
            torch.cuda.synchronize()
            resident_ms = (time.perf_counter() - t0) * 1e3
        out["resident_ms"] = resident_ms
        g2_s, se_s = [np.array(x, copy=True) for x in streamed()]
        cor_s = c.correlators(sample)
        out["parity"] = {"g2_bit_identical": same(g2_r, g2_s), "stderr_bit_identical": same(se_r, se_s),
                         "sampled_pixels": int(sample.size),
                         "G2_IP_IF_bit_identical": [same(x, y) for x, y in zip(cor_r, cor_s)]}
        c.kernel_timing(True)
        c.kernel_report(reset=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        streamed()
        torch.cuda.synchronize()
        out["streamed_ms"] = (time.perf_counter() - t0) * 1e3
        rep = c.kernel_report(reset=True)
        c.kernel_timing(False)
        out["kernels_ms"] = {k: round(v[0], 3) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0]) if v[1] > 0}
        out["kernel_launches"] = {k: int(v[1]) for k, v in rep.items() if v[1] > 0}
        if "k_stream_chunk" in rep:
            ms, n = rep["k_stream_chunk"]
            out["k_stream_chunk_ms_per_chunk"] = ms / max(n, 1)
            out["k_stream_chunk_row_chunks_per_s"] = out["rows"] * n / (ms * 1e-3) if ms > 0 else None
            out["k_stream_chunk_row_frames_per_s"] = out["rows"] * float(F) / (ms * 1e-3) if ms > 0 else None
        t0 = time.perf_counter()
        streamed()
        torch.cuda.synchronize()
        out["streamed_ms_untimed_kernels"] = (time.perf_counter() - t0) * 1e3
        out["frames_per_s_streamed"] = F / (out["streamed_ms_untimed_kernels"] * 1e-3)
        if a.host:
            # the same stream from host memory: every chunk crosses PCIe inside the timed region (one push for the whole job,
            # the library cuts it into chunks and copies chunk k + 1 under k_stream_chunk of chunk k)
            if "h_idx" not in locals():
                h_idx = torch.empty(E, dtype=torch.int32, pin_memory=True).copy_(d_idx)
                h_val = torch.empty(E, dtype=torch.int16, pin_memory=True).copy_(d_val)
                h_off = torch.empty(F + 1, dtype=torch.int64, pin_memory=True).copy_(d_off)
                torch.cuda.synchronize()

            def streamed_host():
                c.reset()
                c.stream_begin(chunk)
                c._check(c._lib.xpcs_stream_push_sparse(c._h, h_idx.data_ptr(), h_val.data_ptr(), h_off.data_ptr(), None, None, F))
                c.stream_finish(want=True)
                c.multitau(want=False)
                return c.normalize()

            g2_h, se_h = [np.array(x, copy=True) for x in streamed_host()]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            streamed_host()
            torch.cuda.synchronize()
            out["streamed_host_ms"] = (time.perf_counter() - t0) * 1e3
            out["h2d_bytes"] = int(E * 6 + (F + 1) * 8)
            out["frames_per_s_streamed_host"] = F / (out["streamed_host_ms"] * 1e-3)
            out["parity"]["host_stream_g2_bit_identical"] = same(g2_r, g2_h)
            ok = ok and out["parity"]["host_stream_g2_bit_identical"]
        ok = ok and out["parity"]["g2_bit_identical"] and all(out["parity"]["G2_IP_IF_bit_identical"])
        lines.append(json.dumps(out))
        print(lines[-1], flush=True)
    c.close()
    if a.out and lines:
        with open(a.out, "w") as f:
            f.write("\n".join(lines) + "\n")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
