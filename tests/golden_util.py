"""Shared loader for tests/golden/*.npz (fixtures written by tests/golden/make_golden.py from the
unmodified reference binary)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(kind=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        n = os.path.splitext(os.path.basename(p))[0]
        if kind is None or str(np.load(p)["in_kind"]) == kind:
            out.append(n)
    return out


class Case:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.kind = str(z["in_kind"])
        self.inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
        self.ref = {k[4:].replace("__", "/"): z[k] for k in z.files if k.startswith("ref_")}
        p = [int(x) for x in self.inp["params"]]
        self.F_raw, self.dpl, self.stride, self.avg, self.swindow, self.norm = p[:6]
        self.darks = p[6] if len(p) > 6 else 0
        block = self.stride * self.avg if (self.stride > 1 and self.avg > 1) else max(self.stride, self.avg)
        self.F = self.F_raw // block
        self.dq, self.sq = self.inp["dq"], self.inp["sq"]
        self.flat = self.inp.get("flat")
        self.fmt = str(self.inp["fmt"]) if "fmt" in self.inp else "imm"   # imm | ufxc | rigaku
        self.late_window = self.fmt == "rigaku"                           # XPCS_COMPAT_LATE_WINDOW
        self.method = str(self.inp["method"]) if "method" in self.inp else "symmetric"   # two-time smoothing


def rel_err(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b)
    d[both_nan] = 0.0
    den = np.maximum(np.abs(b), 1e-300)
    r = d / den
    r[(d == 0)] = 0.0
    return float(np.nanmax(r)) if r.size else 0.0, int(np.sum(np.isnan(a) != np.isnan(b)))


def n_diff(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    return int(np.sum(~((a == b) | (np.isnan(a) & np.isnan(b)))))
