"""What-if timing with the developer build (scripts/build_dbg.sh): which phase of the fused
kernel costs what.  knobs: 1 = no sub-pass barriers, 2 = no global load/store, 4 = no rotation
math, 8 = no register-block processing at all, 16 = no tile store, 32 = no tile load.
env: FFB_KNOBS=0,8,24 (knob sets to time), FFB_SIDE=beta."""
import ctypes
import json
import os
import sys

import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "build", "dbg", "libffsim_b200.so")
if not os.path.exists(lib):
    subprocess.run(["bash", os.path.join(ROOT, "scripts", "build_dbg.sh")], check=True, stdout=subprocess.DEVNULL)
os.environ["FFSIM_B200_LIB"] = lib
sys.path.insert(0, ROOT)
import numpy as np
import torch

import ffsim_b200 as ffsim
from ffsim_b200 import _lib

norb, nelec = int(sys.argv[1]), (int(sys.argv[2]), int(sys.argv[3]))
for kv in filter(None, (sys.argv[4] if len(sys.argv) > 4 else "").split(",")):
    k, v = kv.split("=")
    _lib.set_option(k, int(v))
_lib.lib.ffb_debug_knobs.argtypes = [ctypes.c_int]
u = ffsim.random.random_unitary(norb, seed=1)
vec = torch.randn(ffsim.dim(norb, nelec), dtype=torch.complex128, device="cuda")
res = {}
KNOBS = [int(k) for k in os.environ.get("FFB_KNOBS", "0,2,4,8,6,10").split(",")]
SIDE = (None, u) if os.environ.get("FFB_SIDE") == "beta" else (u, None)
for knobs in KNOBS:
    _lib.lib.ffb_debug_knobs(knobs)
    ts = []
    for it in range(6):
        vec.normal_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ffsim.apply_orbital_rotation(vec, SIDE, norb, nelec, copy=False)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res[knobs] = round(float(np.median(ts[2:])), 4)
print(json.dumps(res))
