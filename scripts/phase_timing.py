"""Per-phase warp-cycle breakdown of fused_pass_kernel (developer build with -DFFB_DEBUG_TIMING).
usage: python scripts/phase_timing.py norb na nb [opts]"""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "build", "dbgt", "libffsim_b200.so")
if not os.path.exists(lib):
    subprocess.run(["bash", os.path.join(ROOT, "scripts", "build_dbg.sh")], check=True,
                   env=dict(os.environ, FFB_EXTRA="-DFFB_DEBUG_TIMING", FFB_DBG_DIR="dbgt"), stdout=subprocess.DEVNULL)
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
u = ffsim.random.random_unitary(norb, seed=1)
vec = torch.randn(ffsim.dim(norb, nelec), dtype=torch.complex128, device="cuda")
NAMES = ["tile load", "table stage + barrier", "chunk fetch/bookkeeping", "gather", "runs (math+dispatch)", "scatter",
         "barrier wait", "tile store"]
buf = (ctypes.c_ulonglong * 16)()
for side, mat in (("alpha", (u, None)), ("beta", (None, u))):
    for _ in range(2):
        ffsim.apply_orbital_rotation(vec, mat, norb, nelec, copy=False)
    _lib.lib.ffb_debug_phase_cycles(buf, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ffsim.apply_orbital_rotation(vec, mat, norb, nelec, copy=False)
    b.record()
    torch.cuda.synchronize()
    _lib.lib.ffb_debug_phase_cycles(buf, 1)
    tot = sum(buf[:8])
    print(json.dumps({"side": side, "ms": round(a.elapsed_time(b), 3),
                      "phases_pct": {n: round(100.0 * buf[i] / tot, 1) for i, n in enumerate(NAMES)}}))
