"""Orbital rotation: same signature as python/ffsim/gates/orbital_rotation.py:44-51.

The reference decomposes the unitary into Givens rotations and calls one Rust
kernel per rotation (orbital_rotation.py:130-151).  Here the same decomposition
(host, csrc/tables.cpp) feeds a plan that fuses all rotations and phase shifts of
both spin sectors into a few passes of one CUDA kernel (csrc/givens_kernels.cu).
"""

from __future__ import annotations

import ctypes
import numbers
from collections import OrderedDict
from typing import Any

import numpy as np
import torch

from ffsim_b200 import _device, _lib
from ffsim_b200.cistring import get_tables, one_subspace_indices, zero_one_subspace_indices
from ffsim_b200.linalg.givens import _decompose_raw

_PLAN_CACHE: "OrderedDict[tuple, _Plan]" = OrderedDict()
_PLAN_CACHE_SIZE = 32


class _Plan:
    def __init__(self, handle, tables):
        self.handle = handle
        self.tables = tables  # keep the table handles alive

    def workspace_bytes(self) -> int:
        return int(_lib.lib.ffb_plan_workspace_bytes(self.handle, 0))

    def describe(self) -> str:
        buf = ctypes.create_string_buffer(4096)
        _lib.check(_lib.lib.ffb_plan_describe(self.handle, buf, len(buf)))
        return buf.value.decode()

    def n_state_passes(self) -> int:
        return int(_lib.lib.ffb_plan_n_state_passes(self.handle))

    def __del__(self):
        try:
            _lib.lib.ffb_plan_destroy(self.handle)
        except Exception:
            pass


def _split_mat(mat: Any) -> tuple[np.ndarray | None, np.ndarray | None]:
    """``_get_givens_decomposition`` argument handling (orbital_rotation.py:157-174)."""
    if isinstance(mat, torch.Tensor):
        mat = mat.detach().cpu().numpy()
    if isinstance(mat, np.ndarray) and mat.ndim == 2:
        return mat, mat
    mat_a, mat_b = mat
    mat_a = None if mat_a is None else _device.as_host_matrix(mat_a)
    mat_b = None if mat_b is None else _device.as_host_matrix(mat_b)
    return mat_a, mat_b


def _side_args(decomp):
    if decomp is None:
        return None, -1, None
    rots, phases = decomp
    return _lib.ptr(rots), len(rots), _lib.ptr(phases)


def get_plan(norb: int, nelec: tuple[int, int], mat_a, mat_b) -> _Plan:
    """Plan for rotating alpha by ``mat_a`` and beta by ``mat_b`` (either may be None)."""
    _device.require_cuda()
    dev = _device.sync_device()
    n_alpha, n_beta = nelec
    same = mat_a is mat_b
    decomp_a = None if mat_a is None else _decompose_raw(mat_a)
    decomp_b = decomp_a if same else (None if mat_b is None else _decompose_raw(mat_b))

    def sig(d):
        return None if d is None else (d[0]["i"].tobytes(), d[0]["j"].tobytes())

    key = (dev, norb, n_alpha, n_beta, sig(decomp_a), sig(decomp_b),
           tuple(_lib.get_option(k) for k in ("smem_bytes", "min_cols", "max_cols", "sub_window",
                                              "threads", "beta_mode", "bulk_copies")))
    ra, na, pa = _side_args(decomp_a)
    rb, nb, pb = _side_args(decomp_b)
    plan = _PLAN_CACHE.get(key)
    if plan is not None:
        _PLAN_CACHE.move_to_end(key)
        _lib.check(_lib.lib.ffb_plan_update_coefficients(plan.handle, ra, na, pa, rb, nb, pb))
        return plan
    ta, tb = get_tables(norb, n_alpha), get_tables(norb, n_beta)
    handle = ctypes.c_void_p()
    _lib.check(
        _lib.lib.ffb_plan_orbital_rotation(ta.handle, tb.handle, ra, na, pa, rb, nb, pb, ctypes.byref(handle))
    )
    plan = _Plan(handle, (ta, tb))
    _PLAN_CACHE[key] = plan
    while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
        _PLAN_CACHE.popitem(last=False)
    return plan


def _rotate_device(t, mat_a, mat_b, norb: int, nelec: tuple[int, int]) -> None:
    """Rotate the device vector ``t`` in place (a CUDA tensor or a ShardedVector)."""
    if _device.is_sharded(t):
        from ffsim_b200 import distributed

        distributed.rotate(t, mat_a, mat_b)
        return
    with torch.cuda.device(t.device):
        plan = get_plan(norb, nelec, mat_a, mat_b)
        ws_bytes = plan.workspace_bytes()
        ws = torch.empty(ws_bytes // 16, dtype=torch.complex128, device=t.device) if ws_bytes else None
        _lib.check(
            _lib.lib.ffb_apply_orbital_rotation(
                plan.handle, t.data_ptr(), ws.data_ptr() if ws is not None else None, _device.stream_ptr()
            )
        )


def _check_dim(t, norb: int, nelec) -> None:
    from ffsim_b200.states import dim

    if _device.is_sharded(t) and (t.norb != norb or t.nelec != tuple(nelec)):
        raise ValueError(f"sharded vector was built for norb={t.norb}, nelec={t.nelec}")
    d = dim(norb, nelec)
    if t.numel() != d:
        raise ValueError(f"vec has {t.numel()} entries, expected {d} for norb={norb}, nelec={nelec}")


def apply_orbital_rotation(
    vec,
    mat,
    norb: int,
    nelec: int | tuple[int, int],
    *,
    copy: bool = True,
):
    r"""Apply an orbital rotation to a vector.

    Maps :math:`a^\dagger_{i\sigma} \mapsto \sum_j U^{(\sigma)}_{ji} a^\dagger_{j\sigma}`.
    Arguments and ``copy`` semantics are those of ``ffsim.apply_orbital_rotation``;
    ``vec`` may also be a CUDA ``torch.complex128`` tensor, in which case a CUDA
    tensor is returned (the input itself, updated in place, when ``copy=False``).
    """
    t, kind = _device.to_device(vec, copy=copy)
    if isinstance(nelec, numbers.Integral):
        # spinless: (dim, 1) matrix, the rotation acts on the alpha index (orbital_rotation.py:102-114)
        nelec_pair = (int(nelec), 0)
        mat_a, mat_b = _device.as_host_matrix(mat), None
    else:
        nelec_pair = (int(nelec[0]), int(nelec[1]))
        mat_a, mat_b = _split_mat(mat)
    _check_dim(t, norb, nelec_pair)
    _rotate_device(t, mat_a, mat_b, norb, nelec_pair)
    return _device.from_device(t, kind)


# ---------------------------------------------------------------------------------
# _lib-level entry points: one launch per call, the reference's FFI granularity.


def _index_tensor(indices, device) -> torch.Tensor:
    arr = np.array(indices, dtype=np.uint64).view(np.int64)
    return torch.from_numpy(arr).to(device)


def apply_givens_rotation_in_place(vec: torch.Tensor, c: float, s: complex, slice1, slice2) -> None:
    """``_lib.apply_givens_rotation_in_place`` (src/gates/orbital_rotation.rs:20) on a 2-D CUDA tensor."""
    if not (_device.is_device_vector(vec) and vec.dim() == 2 and vec.dtype == torch.complex128):
        raise TypeError("vec must be a 2-D complex128 CUDA tensor")
    if vec.stride(1) != 1:
        raise ValueError("vec must have unit column stride")
    with torch.cuda.device(vec.device):
        _device.sync_device()
        s1, s2 = _index_tensor(slice1, vec.device), _index_tensor(slice2, vec.device)
        _lib.check(
            _lib.lib.ffb_apply_givens_rotation_in_place(
                vec.data_ptr(), vec.shape[0], vec.shape[1], vec.stride(0), float(c), _lib.c128(s),
                s1.data_ptr(), s2.data_ptr(), s1.numel(), _device.stream_ptr(),
            )
        )


def apply_phase_shift_in_place(vec: torch.Tensor, phase: complex, indices) -> None:
    """``_lib.apply_phase_shift_in_place`` (src/gates/phase_shift.rs:18) on a 2-D CUDA tensor."""
    if not (_device.is_device_vector(vec) and vec.dim() == 2 and vec.dtype == torch.complex128):
        raise TypeError("vec must be a 2-D complex128 CUDA tensor")
    if vec.stride(1) != 1:
        raise ValueError("vec must have unit column stride")
    with torch.cuda.device(vec.device):
        _device.sync_device()
        idx = _index_tensor(indices, vec.device)
        _lib.check(
            _lib.lib.ffb_apply_phase_shift_in_place(
                vec.data_ptr(), vec.shape[0], vec.shape[1], vec.stride(0), _lib.c128(phase),
                idx.data_ptr(), idx.numel(), _device.stream_ptr(),
            )
        )


def apply_orbital_rotation_unfused(vec, mat, norb: int, nelec: tuple[int, int]):
    """The reference's own loop structure (orbital_rotation.py:117-154) on the GPU:
    one launch per Givens rotation and per phase shift.  Correctness anchor for the
    fused plan and the baseline the fusion is measured against."""
    from ffsim_b200.linalg.givens import givens_decomposition

    t, kind = _device.to_device(vec, copy=True)
    n_alpha, n_beta = nelec
    ta, tb = get_tables(norb, n_alpha), get_tables(norb, n_beta)
    mat_a, mat_b = _split_mat(mat)
    m = t.view(ta.dim, tb.dim)

    def one_side(mv, decomp_mat, nocc):
        rots, phases = givens_decomposition(decomp_mat)
        for c, s, i, j in rots:
            idx = zero_one_subspace_indices(norb, nocc, (i, j))
            half = len(idx) // 2
            apply_givens_rotation_in_place(mv, c, np.conj(s), idx[:half], idx[half:])
        for i, phase in enumerate(phases):
            apply_phase_shift_in_place(mv, phase, one_subspace_indices(norb, nocc, (i,)))

    if mat_a is not None:
        one_side(m, mat_a, n_alpha)
    if mat_b is not None:
        mt = m.t().contiguous()
        one_side(mt, mat_b, n_beta)
        m = mt.t().contiguous()
    return _device.from_device(m.reshape(-1), kind)
