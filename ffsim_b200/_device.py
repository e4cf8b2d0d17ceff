"""Device plumbing: NumPy / torch in, device tensor through the kernels, same kind out.

PyTorch is used for device memory and streams only; every computation on the
state happens in libffsim_b200.so.
"""

from __future__ import annotations

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


def local_block(t, dim_a: int):
    """(tensor holding the locally stored alpha rows, first row, number of rows)."""
    if is_sharded(t):
        return t.local, t.row0, t.n_rows
    return t, 0, dim_a


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


def _upload(arr: np.ndarray) -> torch.Tensor:
    """Host -> device.  Pinned buffers go straight to the DMA engine; large pageable
    ones are staged through a pinned buffer from torch's caching host allocator."""
    if not arr.flags.writeable:
        arr = arr.copy()
    src = torch.from_numpy(arr)
    if arr.nbytes >= _STAGE_MIN_BYTES and not src.is_pinned():
        stage = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        stage.copy_(src)
        src = stage
    out = torch.empty(src.shape, dtype=src.dtype, device="cuda")
    out.copy_(src, non_blocking=True)
    torch.cuda.current_stream().synchronize()  # the staging buffer may be recycled after this
    return out


def _download(t: torch.Tensor) -> torch.Tensor:
    """Device -> host into pinned memory (cached by torch's host allocator)."""
    if t.numel() * t.element_size() >= _STAGE_MIN_BYTES:
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        out.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out
    return t.cpu()


def from_device(t, kind: Kind):
    if kind.sharded:
        return t
    if kind.numpy:
        return _download(t).numpy()
    if kind.torch_cpu:
        return _download(t)
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
