"""Developer timing: where the end-to-end (host buffers in and out) time of one application goes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ffsim_b200 as ffsim
from ffsim_b200 import _device

norb, nelec = 16, (5, 5)
dim = ffsim.dim(norb, nelec)
rng = np.random.default_rng(1)
u = ffsim.random.random_unitary(norb, seed=rng)
mat = ffsim.random.random_real_symmetric_matrix(norb, seed=rng)
pinned = torch.empty(dim, dtype=torch.complex128, pin_memory=True)
pinned.copy_(torch.from_numpy(ffsim.random.random_state_vector(dim, seed=rng)))
host = pinned.numpy()
print("from_numpy(pinned).is_pinned():", torch.from_numpy(host).is_pinned())


def t(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r


ms, dev = t(lambda: _device.to_device(host, copy=True)[0])
print(f"to_device (H2D 305 MB): {ms:.2f} ms  -> {dim * 16 / ms / 1e6:.1f} GB/s")
ms, _ = t(lambda: ffsim.apply_orbital_rotation(dev, u, norb, nelec, copy=False))
print(f"orbital rotation on device: {ms:.2f} ms")
ms, _ = t(lambda: ffsim.apply_diag_coulomb_evolution(dev, mat, 1.0, norb, nelec, copy=False))
print(f"diag coulomb on device: {ms:.2f} ms")
ms, out = t(lambda: _device.from_device(dev, _device.Kind(numpy=True)))
print(f"from_device (D2H 305 MB): {ms:.2f} ms  -> {dim * 16 / ms / 1e6:.1f} GB/s")
raw = torch.empty(dim, dtype=torch.complex128, device="cuda")
ms, _ = t(lambda: raw.copy_(pinned, non_blocking=True))
print(f"raw pinned H2D copy: {ms:.2f} ms -> {dim * 16 / ms / 1e6:.1f} GB/s")
back = torch.empty(dim, dtype=torch.complex128, pin_memory=True)
ms, _ = t(lambda: back.copy_(raw, non_blocking=True))
print(f"raw pinned D2H copy: {ms:.2f} ms -> {dim * 16 / ms / 1e6:.1f} GB/s")
ms, _ = t(lambda: torch.empty(dim, dtype=torch.complex128, pin_memory=True))
print(f"pinned alloc (cached): {ms:.2f} ms")


def api():
    a = ffsim.apply_orbital_rotation(host, u, norb, nelec)
    return ffsim.apply_diag_coulomb_evolution(a, mat, 1.0, norb, nelec)


ms, _ = t(api, 3)
print(f"public API, numpy in/out, two calls (2 uploads + 2 downloads): {ms:.2f} ms")
