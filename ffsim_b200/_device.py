"""Device plumbing: NumPy / torch in, device tensor through the kernels, same kind out.

PyTorch is used for device memory and streams only; every computation on the
state happens in libffsim_b200.so.
"""

from __future__ import annotations

import os
import threading
import weakref
from typing import Any

import numpy as np
import torch

from ffsim_b200 import _lib


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "ffsim_b200 needs a CUDA device: the state-vector kernels are sm_100a CUDA "
            "and there is no CPU fallback."
        )


def sync_device() -> int:
    """Make torch's current device the C library's current device; return its index."""
    dev = torch.cuda.current_device()
    _lib.check(_lib.lib.ffb_set_device(dev))
    return dev


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class Kind:
    """How the caller handed the vector over, so that it gets the same kind back."""

    __slots__ = ("numpy", "torch_cpu", "device", "sharded")

    def __init__(self, numpy=False, torch_cpu=False, device=None, sharded=False):
        self.numpy = numpy
        self.torch_cpu = torch_cpu
        self.device = device
        self.sharded = sharded


def is_sharded(vec: Any) -> bool:
    from ffsim_b200.distributed import ShardedVector

    return isinstance(vec, ShardedVector)


def local_block(t, dim_a: int, dim_b: int):
    """The locally stored block of the (dim_a x dim_b) state:
    (tensor, first alpha row, rows, first beta column, columns, row stride)."""
    if is_sharded(t):
        return t.block()
    return t, 0, dim_a, 0, dim_b, dim_b


def same_layout(t, out) -> None:
    """Binary operators on sharded vectors need both operands in the same distribution."""
    if is_sharded(t) and is_sharded(out):
        out.set_layout_like(t)


def empty_like(t):
    return t.empty_like() if is_sharded(t) else torch.empty_like(t)


def to_device(vec: Any, *, copy: bool) -> tuple[torch.Tensor, Kind]:
    """1-D complex128 CUDA tensor holding ``vec``.

    NumPy arrays and CPU tensors are uploaded (so the result never aliases the
    input); CUDA tensors are cloned only when ``copy`` is set.
    """
    require_cuda()
    if is_sharded(vec):  # row-sharded multi-GPU state: stays sharded
        return (vec.clone() if copy else vec), Kind(sharded=True)
    if isinstance(vec, torch.Tensor):
        if vec.is_cuda:
            with torch.cuda.device(vec.device):
                t = vec.reshape(-1)
                if t.dtype != torch.complex128:
                    t = t.to(torch.complex128)
                elif copy or not t.is_contiguous():
                    t = t.clone(memory_format=torch.contiguous_format)
            return t, Kind(device=vec.device)
        t = vec.detach().reshape(-1).to(torch.complex128).contiguous()
        return _upload(t.numpy()), Kind(torch_cpu=True)
    arr = np.ascontiguousarray(np.asarray(vec).reshape(-1), dtype=np.complex128)
    return _upload(arr), Kind(numpy=True)


_STAGE_MIN_BYTES = 1 << 20

# ---------------------------------------------------------------------------------
# Pinned host buffers.  Page-locking 300 MB costs ~40 ms, seven times the PCIe transfer it
# is meant to speed up, so pinned buffers are pooled by size: staging buffers for uploads go
# straight back to the pool; the buffer behind a result array goes back when the last NumPy
# view of it is garbage collected.

_POOL: dict[int, list[torch.Tensor]] = {}
_POOL_LOCK = threading.Lock()
_POOL_MAX_BYTES = int(os.environ.get("FFSIM_B200_PINNED_POOL_GB", "48")) << 30  # free pinned memory kept around
_pool_bytes = 0


def _pool_get(nbytes: int) -> torch.Tensor:
    """A pinned uint8 tensor of exactly ``nbytes`` bytes (pooled)."""
    global _pool_bytes
    with _POOL_LOCK:
        free = _POOL.get(nbytes)
        if free:
            _pool_bytes -= nbytes
            return free.pop()
    return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)


def _pool_put(buf: torch.Tensor) -> None:
    global _pool_bytes
    n = buf.numel()
    with _POOL_LOCK:
        if _pool_bytes + n <= _POOL_MAX_BYTES:
            _POOL.setdefault(n, []).append(buf)
            _pool_bytes += n


class _PinnedOwner:
    """Base object of a result array: keeps the pinned buffer alive for every view of the array
    and hands it back to the pool when the last of them is collected."""

    __slots__ = ("__array_interface__", "__weakref__")

    def __init__(self, buf: torch.Tensor, n_complex: int):
        self.__array_interface__ = {
            "shape": (n_complex,), "typestr": "<c16", "data": (buf.data_ptr(), False), "version": 3,
        }
        weakref.finalize(self, _pool_put, buf)


def _upload(arr: np.ndarray) -> torch.Tensor:
    """Host -> device.  Pinned buffers go straight to the DMA engine; large pageable ones are
    copied through a pooled pinned staging buffer."""
    src = torch.from_numpy(arr) if arr.flags.writeable else torch.from_numpy(arr.copy())
    out = torch.empty(src.shape, dtype=src.dtype, device="cuda")
    if arr.nbytes >= _STAGE_MIN_BYTES and not src.is_pinned():
        stage = _pool_get(arr.nbytes)
        view = stage.view(torch.complex128)
        view.copy_(src)
        out.copy_(view, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        _pool_put(stage)
        return out
    out.copy_(src, non_blocking=True)
    torch.cuda.current_stream().synchronize()  # the caller may overwrite its buffer after this
    return out


def _download(t: torch.Tensor) -> np.ndarray:
    """Device -> host: a NumPy array over a pooled pinned buffer."""
    nbytes = t.numel() * t.element_size()
    if nbytes >= _STAGE_MIN_BYTES:
        buf = _pool_get(nbytes)
        buf.view(torch.complex128).copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return np.asarray(_PinnedOwner(buf, t.numel()))
    return t.cpu().numpy()


def from_device(t, kind: Kind):
    if kind.sharded:
        return t
    if kind.numpy:
        return _download(t)
    if kind.torch_cpu:
        return torch.from_numpy(_download(t))
    return t


def new_like(t: torch.Tensor) -> torch.Tensor:
    return torch.empty_like(t)


def is_device_vector(vec: Any) -> bool:
    return isinstance(vec, torch.Tensor) and vec.is_cuda


def as_host_matrix(mat: Any, dtype=None) -> np.ndarray:
    """Small operator matrices live on the host (norb x norb)."""
    if isinstance(mat, torch.Tensor):
        mat = mat.detach().cpu().numpy()
    return np.asarray(mat) if dtype is None else np.asarray(mat, dtype=dtype)
