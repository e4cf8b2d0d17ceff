"""Developer probe of the host pipeline: where does the time of overlapped applications go?
usage: python scripts/e2e_probe.py [norb na nb]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ffsim_b200 as ffsim

norb, nelec = (int(sys.argv[1]), (int(sys.argv[2]), int(sys.argv[3]))) if len(sys.argv) > 3 else (16, (5, 5))
dim = ffsim.dim(norb, nelec)
rng = np.random.default_rng(0)
u = ffsim.random.random_unitary(norb, seed=rng)
mat = ffsim.random.random_real_symmetric_matrix(norb, seed=rng)
ops = [("orbital_rotation", u), ("diag_coulomb", mat, 0.5)]
host = ffsim.pinned_empty(dim)
host[:] = 1.0 / np.sqrt(dim)
N = 8

# raw PCIe: one direction at a time, then both at once
dev_a = torch.empty(dim, dtype=torch.complex128, device="cuda")
dev_b = torch.empty(dim, dtype=torch.complex128, device="cuda")
pin_a = torch.empty(dim, dtype=torch.complex128, pin_memory=True)
pin_b = torch.empty(dim, dtype=torch.complex128, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1):
        dev_a.copy_(pin_a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2):
        pin_b.copy_(dev_b, non_blocking=True)
def both():
    h2d(); d2h()
gb = dim * 16 / 1e9
print(f"state {gb:.2f} GB: h2d {t(h2d):.2f} ms, d2h {t(d2h):.2f} ms, both at once {t(both):.2f} ms")
del dev_a, dev_b, pin_a, pin_b

for n_chunks in (None, 1):
    for _ in range(3):
        ffsim.evolve_host(host, ops, norb, nelec, n_chunks=n_chunks)
    t0 = time.perf_counter()
    for _ in range(N):
        r = ffsim.evolve_host(host, ops, norb, nelec, n_chunks=n_chunks)
    seq = (time.perf_counter() - t0) / N * 1e3
    # overlapped
    pend = [ffsim.evolve_host_async(host, ops, norb, nelec, n_chunks=n_chunks) for _ in range(3)]
    for h in pend:
        h.result()
    torch.cuda.synchronize()
    enq, wait = [], []
    pend, last = [], None
    t0 = time.perf_counter()
    for _ in range(N):
        a = time.perf_counter()
        pend.append(ffsim.evolve_host_async(host, ops, norb, nelec, n_chunks=n_chunks))
        b = time.perf_counter()
        enq.append((b - a) * 1e3)
        if len(pend) > 2:
            last = pend.pop(0).result()
            wait.append((time.perf_counter() - b) * 1e3)
    for h in pend:
        last = h.result()
    ovl = (time.perf_counter() - t0) / N * 1e3
    print(f"n_chunks={n_chunks}: one at a time {seq:.2f} ms/step, overlapped {ovl:.2f} ms/step; enqueue ms {np.round(enq, 2).tolist()}; wait ms {np.round(wait, 2).tolist()}")
