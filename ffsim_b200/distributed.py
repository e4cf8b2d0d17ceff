"""Row-sharded state vectors: one process per GPU, alpha strings split across ranks.

Layout: rank r holds the contiguous block of alpha rows
[a_off[r], a_off[r+1]) of the (dim_a x dim_b) state, all beta columns.  Then

* beta-side Givens rotations, diagonal Coulomb / number-operator evolutions and
  contractions are rank-local (the kernels take the row offset of the block);
* alpha-side rotations couple rows of different ranks: the state is redistributed
  to column shards [dim_a x (b_off[r+1]-b_off[r])] with one all-to-all over
  NVLink (NCCL), rotated locally along the row index, and redistributed back;
* scalars (vdot, norm) are a local reduction plus an all_reduce.

A ``ShardedVector`` can be passed wherever the public functions take ``vec``
(``apply_orbital_rotation``, ``apply_diag_coulomb_evolution``, ``apply_unitary``,
``linear_operator(...) @ vec``, ...).  The reference has no counterpart: it is a
single-process NumPy code.

Two implementations of the redistribution exist.  On one NVLink/NVSwitch box the shards live
in symmetric (peer-mapped) memory and ``ffb_exchange_blocks`` stores every block straight into its
final place in the destination GPU's buffer: one kernel instead of pack + all-to-all + unpack, no
send/receive staging buffers.  Everywhere else (``gloo`` on CPU tensors -- which is how the host-side
logic is tested without GPUs --, more than two ranks unless ``FFSIM_B200_EXCHANGE=p2p``, or
``FFSIM_B200_EXCHANGE=nccl``) it is ``all_to_all_single``.
"""

from __future__ import annotations

import ctypes
import math
import os
import weakref
from typing import Sequence

import numpy as np
import torch
import torch.distributed as dist


def partition(n: int, world: int) -> list[int]:
    """Offsets of an as-even-as-possible contiguous split of ``range(n)`` into ``world`` parts."""
    base, extra = divmod(n, world)
    offs = [0]
    for r in range(world):
        offs.append(offs[-1] + base + (1 if r < extra else 0))
    return offs


class ShardedVector:
    """A state vector whose alpha rows are distributed over the ranks of ``group``."""

    def __init__(self, local: torch.Tensor, norb: int, nelec: tuple[int, int], group=None):
        self.norb = int(norb)
        self.nelec = (int(nelec[0]), int(nelec[1]))
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dim_a = math.comb(self.norb, self.nelec[0])
        self.dim_b = math.comb(self.norb, self.nelec[1])
        self.a_off = partition(self.dim_a, self.world)
        self.b_off = partition(self.dim_b, self.world)
        self.row0 = self.a_off[self.rank]
        self.n_rows = self.a_off[self.rank + 1] - self.row0
        local = local.reshape(-1)
        if local.dtype != torch.complex128:
            local = local.to(torch.complex128)
        if local.numel() != self.n_rows * self.dim_b:
            raise ValueError(
                f"local block has {local.numel()} entries, expected {self.n_rows} x {self.dim_b}"
            )
        self.local = local.contiguous()

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_global(cls, vec, norb, nelec, group=None, device=None) -> "ShardedVector":
        """Every rank passes the same full vector and keeps its rows (small cases / tests)."""
        full = torch.as_tensor(np.asarray(vec) if not isinstance(vec, torch.Tensor) else vec)
        dim_b = math.comb(norb, nelec[1])
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        offs = partition(math.comb(norb, nelec[0]), world)
        local = full.reshape(-1, dim_b)[offs[rank] : offs[rank + 1]].to(torch.complex128)
        if device is not None:
            local = local.to(device)
        return cls(local.clone(), norb, nelec, group)

    @classmethod
    def hartree_fock(cls, norb, nelec, group=None, device="cuda") -> "ShardedVector":
        """``hartree_fock_state`` (one-hot at address 0), created shard by shard."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        offs = partition(math.comb(norb, nelec[0]), world)
        n_rows = offs[rank + 1] - offs[rank]
        local = torch.zeros(n_rows * math.comb(norb, nelec[1]), dtype=torch.complex128, device=device)
        if offs[rank] == 0 and n_rows > 0:
            local[0] = 1
        return cls(local, norb, nelec, group)

    # ------------------------------------------------------------------ tensor-like helpers
    @property
    def device(self):
        return self.local.device

    def numel(self) -> int:
        return self.dim_a * self.dim_b

    def clone(self) -> "ShardedVector":
        return ShardedVector(self.local.clone(), self.norb, self.nelec, self.group)

    def empty_like(self) -> "ShardedVector":
        return ShardedVector(torch.empty_like(self.local), self.norb, self.nelec, self.group)

    def copy_(self, other: "ShardedVector") -> "ShardedVector":
        self.local.copy_(other.local)
        return self

    def gather(self) -> torch.Tensor:
        """The full vector on every rank (tests / small cases only)."""
        if self.world == 1:
            return self.local.clone()
        # all_gather wants equal sizes: pad every block to the largest one
        n_max = max(self.a_off[r + 1] - self.a_off[r] for r in range(self.world)) * self.dim_b
        mine = torch.zeros(n_max, dtype=torch.complex128, device=self.device)
        mine[: self.local.numel()] = self.local
        parts = [torch.empty(n_max, dtype=torch.complex128, device=self.device) for _ in range(self.world)]
        dist.all_gather([torch.view_as_real(p) for p in parts], torch.view_as_real(mine), group=self.group)
        return torch.cat(
            [parts[r][: (self.a_off[r + 1] - self.a_off[r]) * self.dim_b] for r in range(self.world)]
        )

    def vdot(self, other: "ShardedVector") -> complex:
        """<self|other> (conjugate-linear in self), reduced over the ranks."""
        if self.local.is_cuda:
            # the library's deterministic two-stage reduction (64-bit length: a C5 shard has 2.3e9
            # amplitudes, more than cuBLAS's 32-bit dot accepts)
            from ffsim_b200 import _device, _lib

            acc = torch.zeros(1, 2, dtype=torch.float64, device=self.local.device)
            with torch.cuda.device(self.local.device):
                _device.sync_device()
                _lib.check(_lib.lib.ffb_vdot(self.local.data_ptr(), other.local.data_ptr(), self.local.numel(),
                                             acc.data_ptr(), _device.stream_ptr()))
        else:  # host tensors: the gloo tests of the redistribution logic
            acc = torch.view_as_real(torch.vdot(self.local, other.local).reshape(1)).clone()
        if self.world > 1:
            dist.all_reduce(acc, group=self.group)
        return complex(acc[0, 0].item(), acc[0, 1].item())

    def norm(self) -> float:
        return math.sqrt(max(self.vdot(self).real, 0.0))


# ---------------------------------------------------------------------- redistribution

def to_column_shards(sv: ShardedVector, release: bool = False) -> torch.Tensor:
    """All-to-all #1: row shards [n_rows x dim_b] -> column shards [dim_a x n_cols_local].

    Each rank sends, to rank d, its rows restricted to d's beta columns; what it receives
    from rank s are s's rows restricted to its own columns, and concatenating the sources
    in rank order is exactly the row-major [dim_a x n_cols_local] matrix.

    With ``release`` the row shard's storage is dropped as soon as it has been packed
    (``sv.local`` becomes None until ``from_column_shards`` rebuilds it), which keeps the
    peak at two shard-sized buffers.
    """
    w, r = sv.world, sv.rank
    nb_local = sv.b_off[r + 1] - sv.b_off[r]
    local2d = sv.local.view(sv.n_rows, sv.dim_b)
    if w == 1:
        return local2d if release else local2d.clone()
    send = torch.empty(sv.n_rows * sv.dim_b, dtype=torch.complex128, device=sv.device)
    in_splits, pos = [], 0
    for d in range(w):
        nb = sv.b_off[d + 1] - sv.b_off[d]
        n = sv.n_rows * nb
        send[pos : pos + n].view(sv.n_rows, nb).copy_(local2d[:, sv.b_off[d] : sv.b_off[d + 1]])
        in_splits.append(n)
        pos += n
    device = sv.device
    if release:
        del local2d
        sv.local = None
    out_splits = [(sv.a_off[s + 1] - sv.a_off[s]) * nb_local for s in range(w)]
    recv = torch.empty(sv.dim_a * nb_local, dtype=torch.complex128, device=device)
    dist.all_to_all_single(
        torch.view_as_real(recv), torch.view_as_real(send),
        output_split_sizes=out_splits, input_split_sizes=in_splits, group=sv.group,
    )
    return recv.view(sv.dim_a, nb_local)


def from_column_shards(sv: ShardedVector, cols: torch.Tensor) -> None:
    """All-to-all #2: column shards back into ``sv.local`` (row shards).

    When ``sv.local`` was released, ``cols`` is consumed: pass the only reference to it.
    """
    w, r = sv.world, sv.rank
    nb_local = sv.b_off[r + 1] - sv.b_off[r]
    device = cols.device
    if w == 1:
        if sv.local is None:
            sv.local = cols.reshape(-1)
        elif sv.local.data_ptr() != cols.data_ptr():
            sv.local.view(sv.n_rows, sv.dim_b).copy_(cols)
        return
    send = cols.reshape(-1)  # rows of destination d are contiguous
    in_splits = [(sv.a_off[d + 1] - sv.a_off[d]) * nb_local for d in range(w)]
    out_splits = [sv.n_rows * (sv.b_off[s + 1] - sv.b_off[s]) for s in range(w)]
    recv = torch.empty(sv.n_rows * sv.dim_b, dtype=torch.complex128, device=device)
    dist.all_to_all_single(
        torch.view_as_real(recv), torch.view_as_real(send),
        output_split_sizes=out_splits, input_split_sizes=in_splits, group=sv.group,
    )
    del send
    if sv.local is None:
        cols.untyped_storage().resize_(0)  # hand the column shard's memory back before unpacking
        sv.local = torch.empty(sv.n_rows * sv.dim_b, dtype=torch.complex128, device=device)
    local2d = sv.local.view(sv.n_rows, sv.dim_b)
    pos = 0
    for s in range(w):
        nb = sv.b_off[s + 1] - sv.b_off[s]
        n = sv.n_rows * nb
        local2d[:, sv.b_off[s] : sv.b_off[s + 1]].copy_(recv[pos : pos + n].view(sv.n_rows, nb))
        pos += n


def all_to_all_bytes(sv: ShardedVector) -> int:
    """Bytes this rank sends over NVLink in one redistribution (its off-rank share of the shard)."""
    own = sv.b_off[sv.rank + 1] - sv.b_off[sv.rank]
    return 16 * sv.n_rows * (sv.dim_b - own)


# ---------------------------------------------------------------------- peer-memory exchange

class _SymmBuffer:
    """A symmetric-memory buffer: the same allocation made by every rank, each mapped into all others."""

    def __init__(self, numel: int, device, group):
        import torch.distributed._symmetric_memory as symm_mem

        pg = group if group is not None else dist.group.WORLD
        try:  # needed by some torch versions, a deprecated no-op in others
            symm_mem.enable_symm_mem_for_group(pg.group_name)
        except Exception:
            pass
        self.tensor = symm_mem.empty(numel, dtype=torch.complex128, device=device)
        self.handle = symm_mem.rendezvous(self.tensor, pg)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.numel = numel


_SYMM_POOL: dict[tuple, list[_SymmBuffer]] = {}
_SYMM_STATE = {"checked": False, "ok": False}


def _symm_get(numel: int, device, group) -> _SymmBuffer:
    """Pooled: symmetric allocations are collective and slow, and every rank takes and returns them at
    the same points of the (SPMD) program, so the pools stay identical across ranks."""
    free = _SYMM_POOL.setdefault((id(group), str(device), numel), [])
    return free.pop() if free else _SymmBuffer(numel, device, group)


def _symm_put(buf: _SymmBuffer, device, group) -> None:
    _SYMM_POOL.setdefault((id(group), str(device), buf.numel), []).append(buf)


def p2p_available(sv: "ShardedVector") -> bool:
    """Peer-memory exchange needs CUDA shards, more than one rank, at most 16 of them on one box with
    symmetric memory working; the decision is taken collectively so that all ranks agree."""
    if sv.world == 1 or sv.world > 16 or not sv.device.type == "cuda":
        return False
    # FFSIM_B200_EXCHANGE: "nccl" = never, "p2p" = whenever symmetric memory works, "auto" (default) =
    # only in the configuration it has been validated in (two ranks).  An 8-rank run of the 254 GB state
    # with 32 GB symmetric buffers did not complete within the time limit of the last GPU slot of round 1,
    # so beyond two ranks the NCCL all-to-all (measured at 8 ranks) stays the default until that is understood.
    mode = os.environ.get("FFSIM_B200_EXCHANGE", "auto").lower()
    if mode == "nccl" or (mode != "p2p" and sv.world > 2):
        return False
    if not _SYMM_STATE["checked"]:
        ok = 1
        try:
            probe = _SymmBuffer(16, sv.device, sv.group)
            ok = int(len(probe.ptrs) == sv.world and all(probe.ptrs))
        except Exception:
            ok = 0
        flag = torch.tensor([ok], device=sv.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=sv.group)
        _SYMM_STATE["ok"] = bool(flag.item())
        _SYMM_STATE["checked"] = True
    return _SYMM_STATE["ok"]


def _i64(values):
    return (ctypes.c_int64 * len(values))(*[int(v) for v in values])


def _exchange(src: torch.Tensor, src_ld: int, rows, width, src_off, dst_ptrs, dst_off, dst_ld) -> None:
    from ffsim_b200 import _device, _lib

    n = len(rows)
    ptrs = (ctypes.c_void_p * n)(*[ctypes.c_void_p(int(p)) for p in dst_ptrs])
    _lib.check(_lib.lib.ffb_exchange_blocks(
        src.data_ptr(), int(src_ld), n, _i64(rows), _i64(width), _i64(src_off), ptrs, _i64(dst_off), _i64(dst_ld),
        _device.stream_ptr()))


def _make_symmetric(sv: "ShardedVector") -> _SymmBuffer:
    """Move the row shard into symmetric memory (once per vector; it stays there)."""
    buf = getattr(sv, "_symm", None)
    if buf is not None and sv.local is not None and sv.local.data_ptr() == buf.tensor.data_ptr():
        return buf
    numel = max(max(sv.a_off[r + 1] - sv.a_off[r] for r in range(sv.world)) * sv.dim_b, 1)
    buf = _symm_get(numel, sv.device, sv.group)
    view = buf.tensor[: sv.n_rows * sv.dim_b]
    view.copy_(sv.local)
    sv.local = view
    sv._symm = buf
    weakref.finalize(sv, _symm_put, buf, sv.device, sv.group)
    return buf


def rotate_alpha_p2p(sv: "ShardedVector", plan, stream) -> None:
    """Alpha-side rotation through peer memory: scatter the row shard into every rank's column shard,
    rotate locally, scatter back.  Two kernels move data; nothing is packed, staged or unpacked."""
    from ffsim_b200 import _lib

    w, r = sv.world, sv.rank
    rows_of = [sv.a_off[d + 1] - sv.a_off[d] for d in range(w)]
    cols_of = [sv.b_off[d + 1] - sv.b_off[d] for d in range(w)]
    nb_local = cols_of[r]
    row_buf = _make_symmetric(sv)
    col_numel = max(sv.dim_a * max(cols_of), 1)
    col_buf = _symm_get(col_numel, sv.device, sv.group)
    try:
        # every rank may still be reading its column buffer from an earlier exchange
        col_buf.handle.barrier(channel=0)
        # row shard -> column shards: block d = my rows x d's columns, lands at my row offset of d's buffer
        _exchange(sv.local, sv.dim_b, [sv.n_rows] * w, cols_of, [sv.b_off[d] for d in range(w)],
                  col_buf.ptrs, [sv.a_off[r] * cols_of[d] for d in range(w)], cols_of)
        col_buf.handle.barrier(channel=0)
        if nb_local > 0:
            _lib.check(_lib.lib.ffb_apply_orbital_rotation_rows(
                plan.handle, 0, col_buf.tensor.data_ptr(), nb_local, nb_local, stream))
        # column shard -> row shards: block d = d's rows x my columns, lands at my column offset of d's shard
        _exchange(col_buf.tensor, nb_local, rows_of, [nb_local] * w, [sv.a_off[d] * nb_local for d in range(w)],
                  row_buf.ptrs, [sv.b_off[r]] * w, [sv.dim_b] * w)
        col_buf.handle.barrier(channel=0)
    finally:
        _symm_put(col_buf, sv.device, sv.group)


# ---------------------------------------------------------------------- device ops on shards

def rotate(sv: ShardedVector, mat_a, mat_b) -> None:
    """Orbital rotation of a sharded state, in place (both spin sectors)."""
    from ffsim_b200 import _device, _lib
    from ffsim_b200.gates.orbital_rotation import get_plan

    with torch.cuda.device(sv.device):
        plan = get_plan(sv.norb, sv.nelec, mat_a, mat_b)
        stream = _device.stream_ptr()
        if mat_b is not None and sv.n_rows > 0:
            if _lib.lib.ffb_plan_beta_in_place(plan.handle):
                _lib.check(_lib.lib.ffb_apply_orbital_rotation_strided(
                    plan.handle, 1, sv.local.data_ptr(), sv.n_rows, 1, sv.dim_b, stream))
            else:
                ws = torch.empty(sv.dim_b * sv.n_rows, dtype=torch.complex128, device=sv.device)
                _lib.check(_lib.lib.ffb_transpose(sv.local.data_ptr(), ws.data_ptr(), sv.n_rows, sv.dim_b,
                                                  sv.dim_b, sv.n_rows, stream))
                _lib.check(_lib.lib.ffb_apply_orbital_rotation_rows(
                    plan.handle, 1, ws.data_ptr(), sv.n_rows, sv.n_rows, stream))
                _lib.check(_lib.lib.ffb_transpose(ws.data_ptr(), sv.local.data_ptr(), sv.dim_b, sv.n_rows,
                                                  sv.n_rows, sv.dim_b, stream))
        if mat_a is not None and p2p_available(sv):
            rotate_alpha_p2p(sv, plan, stream)
        elif mat_a is not None:
            cols = to_column_shards(sv, release=True)
            nb_local = cols.shape[1]
            if nb_local > 0:
                _lib.check(_lib.lib.ffb_apply_orbital_rotation_rows(
                    plan.handle, 0, cols.data_ptr(), nb_local, nb_local, stream))
            from_column_shards(sv, cols)
