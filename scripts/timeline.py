"""Per-warp timeline of fused_pass_kernel's first tile on CTA 0 (developer build with
-DFFB_DEBUG_TIMELINE): which warp is gathering, computing, scattering or waiting when, and how much
of the time each scheduler (warp % 4) has at least one warp in the math phase.

usage: python scripts/timeline.py norb na nb [opts] [--chart]      (written without a GPU at hand: first run pending)
"""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "build", "dbgtl", "libffsim_b200.so")
if not os.path.exists(lib):
    subprocess.run(["bash", os.path.join(ROOT, "scripts", "build_dbg.sh")], check=True,
                   env=dict(os.environ, FFB_EXTRA="-DFFB_DEBUG_TIMELINE", FFB_DBG_DIR="dbgtl"),
                   stdout=subprocess.DEVNULL)
os.environ["FFSIM_B200_LIB"] = lib
sys.path.insert(0, ROOT)
import numpy as np
import torch

import ffsim_b200 as ffsim
from ffsim_b200 import _lib

args = [a for a in sys.argv[1:] if not a.startswith("--")]
norb, nelec = int(args[0]), (int(args[1]), int(args[2]))
for kv in filter(None, (args[3] if len(args) > 3 else "").split(",")):
    k, v = kv.split("=")
    _lib.set_option(k, int(v))
NAMES = ["tile load", "table stage + barrier", "chunk fetch", "gather", "math + dispatch", "scatter", "barrier wait",
         "tile store"]
GLYPH = "L.fgMsbS"
CAP = 2048
rec = np.dtype([("t0", np.uint64), ("t1", np.uint64), ("phase", np.int32), ("pad", np.int32)])
buf = np.zeros((32, CAP), dtype=rec)
counts = (ctypes.c_int * 32)()
fn = _lib.lib.ffb_debug_timeline
fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
u = ffsim.random.random_unitary(norb, seed=1)
vec = torch.randn(ffsim.dim(norb, nelec), dtype=torch.complex128, device="cuda")
for side, mat in (("alpha", (u, None)), ("beta", (None, u))):
    ffsim.apply_orbital_rotation(vec, mat, norb, nelec, copy=False)  # warm-up (tables, plan)
    fn(buf.ctypes.data, counts)  # clear
    ffsim.apply_orbital_rotation(vec, mat, norb, nelec, copy=False)
    assert fn(buf.ctypes.data, counts) == CAP
    n_warps = max(w + 1 for w in range(32) if counts[w]) if any(counts) else 0
    ev = {w: buf[w, : min(counts[w], CAP)] for w in range(n_warps)}
    # a multi-pass rotation logs one tile per launch (each starts with a "tile load" event, and CTA 0 may
    # run on another SM with another clock): keep the launch with the most events (the fullest pass)
    starts = np.flatnonzero(ev[0]["phase"] == 0) if n_warps else []
    if len(starts) > 1:
        bounds = list(starts) + [len(ev[0])]
        best = max(range(len(starts)), key=lambda i: bounds[i + 1] - bounds[i])
        for w in list(ev):
            st = np.flatnonzero(ev[w]["phase"] == 0)
            if len(st) == len(starts):
                b = list(st) + [len(ev[w])]
                ev[w] = ev[w][b[best]:b[best + 1]]
    t_begin = min(int(e["t0"].min()) for e in ev.values() if len(e))
    t_end = max(int(e["t1"].max()) for e in ev.values() if len(e))
    span = t_end - t_begin
    per_phase = np.zeros(8)
    for e in ev.values():
        for ph in range(8):
            m = e["phase"] == ph
            per_phase[ph] += float((e["t1"][m] - e["t0"][m]).sum())
    # scheduler occupancy: fraction of the span with >= 1 warp of the scheduler in the math phase / in gather+scatter
    res = 256
    grid = np.zeros((n_warps, res), dtype=np.int8) - 1
    for w, e in ev.items():
        for r in e:
            a = int((int(r["t0"]) - t_begin) * res / span)
            b = max(a + 1, int((int(r["t1"]) - t_begin) * res / span))
            grid[w, a:min(b, res)] = r["phase"]
    occ = {}
    for name, phases in (("math", (4,)), ("lsu", (3, 5)), ("barrier", (6,))):
        per_sched = []
        for sch in range(4):
            rows = grid[sch::4]
            per_sched.append(float(np.isin(rows, phases).any(axis=0).mean()) if len(rows) else 0.0)
        occ[name] = [round(x, 3) for x in per_sched]
    print(json.dumps({"side": side, "warps": n_warps, "events": int(sum(counts[:n_warps])), "span_cycles": span,
                      "warp_time_pct": {n: round(100 * per_phase[i] / per_phase.sum(), 1) for i, n in enumerate(NAMES)},
                      "scheduler_has_warp_in_phase": occ}))
    if "--chart" in sys.argv:
        print("  one row per warp, %d columns = %d cycles; %s" % (res, span, ", ".join(f"{g}={n}" for g, n in zip(GLYPH, NAMES))))
        for w in range(n_warps):
            print("  w%02d " % w + "".join(GLYPH[p] if p >= 0 else " " for p in grid[w]))
