import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ffsim_b200 import _device as d
n = 19079424
t = torch.randn(n, dtype=torch.complex128, device="cuda")
torch.cuda.synchronize()
def now():
    torch.cuda.synchronize(); return time.perf_counter()
keep = None
for i in range(6):
    t0 = now(); buf = d._pool_get(n * 16); t1 = now()
    v = buf.view(torch.complex128); v.copy_(t, non_blocking=True); torch.cuda.current_stream().synchronize(); t2 = now()
    arr = np.asarray(d._PinnedOwner(buf, n)); t3 = now()
    del buf, v
    keep = arr; t4 = now()
    print(f"iter {i}: pool_get {1e3*(t1-t0):.2f} copy {1e3*(t2-t1):.2f} asarray {1e3*(t3-t2):.2f} rebind(frees previous) {1e3*(t4-t3):.2f} pool={ {k: len(x) for k, x in d._POOL.items()} }")
for i in range(4):
    t0 = now(); r = d._download(t); t1 = now(); keep = r; t2 = now()
    print(f"_download {1e3*(t1-t0):.2f} ms, rebind {1e3*(t2-t1):.2f}")
