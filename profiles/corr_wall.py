"""Wall time of the host program `corr config.hdf5 --imm data.imm` (IMM file in /dev/shm -> result datasets in the
HDF5 file) on a BASELINE configuration, next to the reference's own stage lines for the same files when
oracle/_ref/corr_ref is present.  usage: python profiles/corr_wall.py c1|c3 [gpus]"""
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402
from oracle import refdrv  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "c1"
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pkg = entry.load_package()
wl = bench.WORKLOADS[key]
h, w, F, occ = wl["h"], wl["w"], wl["F"], wl["occ"]
dq, sq = bench.module_maps(pkg, wl, 1)
off, idx, val = pkg.synth.sparse_frames(h * w, F, occ, seed=1234)
d = refdrv.scratch_dir()
imm = os.path.join(d, "data.imm")
pkg.synth.write_imm_sparse(imm, h, w, off, idx, val)
cfg = os.path.join(d, "config.hdf5")
f = pkg.h5lite.File()
for path, value in refdrv.config_items(dq, sq, F, imm, dpl=8)[0]:
    f.put(path, value)
f.save(cfg)
f.close()
corr = os.path.join(os.path.dirname(pkg.cabi.LIB_PATH), "corr")
out = {"workload": wl["name"], "imm_bytes": os.path.getsize(imm), "runs": []}
for it in range(3):
    shutil.copy(cfg, cfg + ".run")
    t0 = time.perf_counter()
    p = subprocess.run([corr, cfg + ".run", "--gpus", str(gpus)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    dt = time.perf_counter() - t0
    assert p.returncode == 0, p.stdout[-2000:]
    scopes = {n.strip(): float(v) * {"ms": 1e-3, "s": 1.0, "m": 60.0}[u] for n, v, u in refdrv._SCOPE.findall(p.stdout)}
    out["runs"].append({"wall_s": dt, "scopes_s": scopes})
out["best_wall_s"] = min(r["wall_s"] for r in out["runs"])
out["frames_per_s"] = F / out["best_wall_s"]
print(json.dumps(out, indent=1))
shutil.rmtree(d, ignore_errors=True)
