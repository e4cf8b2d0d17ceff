"""Sharded state vectors: one process per GPU, one state distributed over all of them.

A ``ShardedVector`` holds its share of the (dim_a x dim_b) state in one of two distributions
(SURVEY.md section 8e):

* ``"rows"``: rank r holds alpha rows [a_off[r], a_off[r+1]) x all beta columns.  Beta-side Givens
  rotations are local;
* ``"cols"``: rank r holds all alpha rows x beta columns [b_off[r], b_off[r+1]).  Alpha-side
  rotations are local.

Diagonal Coulomb / number-operator evolutions and contractions, axpby and dot products are local in
either distribution (the kernels take the row and column offsets of the block).  The state is NOT
moved back after an alpha-side rotation: it stays where the last operation left it, and the next
rotation starts with the spin sector that is local.  An orbital rotation on both spins therefore
costs ONE redistribution (an all-to-all over NVLink), a LUCJ circuit with L layers L + 1 of them.
Scalars (vdot, norm) are a local reduction plus an all_reduce.

A ``ShardedVector`` can be passed wherever the public functions take ``vec``
(``apply_orbital_rotation``, ``apply_diag_coulomb_evolution``, ``apply_unitary``,
``linear_operator(...) @ vec``, ...).  The reference has no counterpart: it is a single-process
NumPy code.

Two implementations of the redistribution exist.  On one NVLink/NVSwitch box the shards live in
symmetric (peer-mapped) memory and ``ffb_exchange_blocks`` stores every block straight into its
final place in the destination GPU's buffer: one kernel, no staging.  Otherwise it is NCCL
``all_to_all_single`` with the strided side packed / unpacked by the same library kernel
(``ffb_exchange_blocks`` with local destinations); on CPU tensors (``gloo``: how the host-side logic is
tested without GPUs) the pack / unpack are plain tensor copies.  ``FFSIM_B200_EXCHANGE`` = ``p2p`` /
``nccl`` forces one of them, ``auto`` (default) takes the peer-memory path up to
``FFSIM_B200_P2P_MAX_WORLD`` ranks (default 8) and ``FFSIM_B200_P2P_MAX_GB`` per symmetric buffer (default 16):
the configurations measured on this pool (2, 4 and 8 ranks of a 16 GB state).
"""

from __future__ import annotations

import ctypes
import math
import os
import weakref

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


ROWS, COLS = "rows", "cols"

# counters of the redistribution (bench.py / tests read them): exchanges done, bytes this rank sent, the
# implementation used ("p2p" / "nccl") and, when "time" is set, a CUDA event pair around every exchange
STATS = {"exchanges": 0, "bytes_sent": 0, "mode": None, "time": False, "events": []}


class ShardedVector:
    """A state vector distributed over the ranks of ``group`` (see the module docstring)."""

    def __init__(self, local: torch.Tensor, norb: int, nelec: tuple[int, int], group=None, layout: str = ROWS):
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
        self.col0 = self.b_off[self.rank]
        self.n_cols = self.b_off[self.rank + 1] - self.col0
        if layout not in (ROWS, COLS):
            raise ValueError(f"layout must be 'rows' or 'cols', got {layout!r}")
        self.layout = layout
        local = local.reshape(-1)
        if local.dtype != torch.complex128:
            local = local.to(torch.complex128)
        if local.numel() != self.local_numel(layout):
            raise ValueError(f"local block has {local.numel()} entries, expected {self.local_numel(layout)} "
                             f"for the '{layout}' distribution")
        self.local = local.contiguous()
        self._symm = None  # the symmetric-memory buffer behind ``local`` (peer-memory path)

    def local_numel(self, layout: str) -> int:
        return self.n_rows * self.dim_b if layout == ROWS else self.dim_a * self.n_cols

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

    def block(self):
        """(tensor, first alpha row, rows, first beta column, columns, row stride) of the local block."""
        if self.layout == ROWS:
            return self.local, self.row0, self.n_rows, 0, self.dim_b, self.dim_b
        return self.local, 0, self.dim_a, self.col0, self.n_cols, self.n_cols

    def clone(self) -> "ShardedVector":
        return ShardedVector(self.local.clone(), self.norb, self.nelec, self.group, self.layout)

    def empty_like(self) -> "ShardedVector":
        return ShardedVector(torch.empty_like(self.local), self.norb, self.nelec, self.group, self.layout)

    def copy_(self, other: "ShardedVector") -> "ShardedVector":
        if other.layout != self.layout:
            self._adopt(torch.empty(self.local_numel(other.layout), dtype=torch.complex128, device=self.device),
                        other.layout)
        self.local.copy_(other.local)
        return self

    def _adopt(self, local: torch.Tensor, layout: str, symm=None) -> None:
        """New storage for the shard.  A symmetric buffer the vector owned goes back to the pool unless it
        is the one being adopted (``_symm_box`` is what the vector's finalizer releases)."""
        box = getattr(self, "_symm_box", None)
        if box is not None and box[0] is not None and box[0] is not symm:
            _symm_put(box[0], self.local.device if self.local is not None else local.device, self.group)
            box[0] = None
        if symm is not None:
            if box is None:
                box = self._symm_box = [None]
                weakref.finalize(self, _symm_release, box, local.device, self.group)
            box[0] = symm
        self.local, self.layout, self._symm = local, layout, symm

    def set_layout(self, layout: str) -> "ShardedVector":
        """Redistribute (collective: every rank calls it) unless already there."""
        if layout != self.layout:
            redistribute(self, layout)
        return self

    def set_layout_like(self, other: "ShardedVector") -> "ShardedVector":
        return self.set_layout(other.layout)

    def gather(self) -> torch.Tensor:
        """The full vector on every rank (tests / small cases only; collective)."""
        self.set_layout(ROWS)
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
        other.set_layout_like(self)
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


# ---------------------------------------------------------------------- strided block copies

def _i64(values):
    return (ctypes.c_int64 * len(values))(*[int(v) for v in values])


def _copy_blocks(src: torch.Tensor, src_ld, rows, width, src_off, dst_ptrs, dst_off, dst_ld) -> None:
    """``ffb_copy_blocks``: block d = rows[d] x width[d] elements from ``src + src_off[d]`` (row stride
    ``src_ld[d]``, or one stride for all blocks) to ``dst_ptrs[d] + dst_off[d]`` (row stride
    ``dst_ld[d]``).  The destinations are device pointers: local buffers (pack / unpack) or peer memory
    (the exchange itself)."""
    from ffsim_b200 import _device, _lib

    n = len(rows)
    if isinstance(src_ld, int):
        src_ld = [src_ld] * n
    ptrs = (ctypes.c_void_p * n)(*[ctypes.c_void_p(int(p)) for p in dst_ptrs])
    _lib.check(_lib.lib.ffb_copy_blocks(
        src.data_ptr(), n, _i64(rows), _i64(width), _i64(src_off), _i64(src_ld), ptrs, _i64(dst_off), _i64(dst_ld),
        _device.stream_ptr()))


def _row_block_geometry(sv: ShardedVector):
    """Per peer d: the block (my rows x d's columns) inside the rows-distribution shard, and where it
    sits in d's cols-distribution shard."""
    w, r = sv.world, sv.rank
    cols_of = [sv.b_off[d + 1] - sv.b_off[d] for d in range(w)]
    return {
        "rows": [sv.n_rows] * w, "width": cols_of,
        "row_off": [sv.b_off[d] for d in range(w)],                 # offset inside my [n_rows x dim_b] shard
        "col_off": [sv.a_off[r] * cols_of[d] for d in range(w)],    # offset inside d's [dim_a x cols_of[d]] shard
        "col_ld": cols_of,
    }


def _col_block_geometry(sv: ShardedVector):
    """Per peer d: the block (d's rows x my columns) inside the cols-distribution shard, and where it
    sits in d's rows-distribution shard."""
    w, r = sv.world, sv.rank
    rows_of = [sv.a_off[d + 1] - sv.a_off[d] for d in range(w)]
    return {
        "rows": rows_of, "width": [sv.n_cols] * w,
        "col_off": [sv.a_off[d] * sv.n_cols for d in range(w)],    # offset inside my [dim_a x n_cols] shard
        "row_off": [sv.b_off[r]] * w,                                # offset inside d's [rows_of[d] x dim_b] shard
        "row_ld": [sv.dim_b] * w,
    }


# ---------------------------------------------------------------------- redistribution (all-to-all)

def all_to_all_bytes(sv: ShardedVector) -> int:
    """Bytes this rank sends over NVLink in one redistribution (its off-rank share of the shard)."""
    if sv.layout == ROWS:
        return 16 * sv.n_rows * (sv.dim_b - sv.n_cols)
    return 16 * (sv.dim_a - sv.n_rows) * sv.n_cols


def _rows_to_cols_a2a(sv: ShardedVector) -> None:
    """rows -> cols through ``all_to_all_single``.  The send side is strided (a column slice of every
    local row per destination) and is packed first; what arrives from rank s are s's rows restricted to
    my columns, and the sources in rank order ARE the row-major [dim_a x n_cols] shard: no unpack."""
    w = sv.world
    g = _row_block_geometry(sv)
    device = sv.device
    send = torch.empty(sv.n_rows * sv.dim_b, dtype=torch.complex128, device=device)
    in_splits = [sv.n_rows * g["width"][d] for d in range(w)]
    pos = [0]
    for n in in_splits:
        pos.append(pos[-1] + n)
    if send.is_cuda:
        with torch.cuda.device(device):
            _copy_blocks(sv.local, sv.dim_b, g["rows"], g["width"], g["row_off"], [send.data_ptr()] * w,
                         pos[:-1], g["width"])
    else:
        local2d = sv.local.view(sv.n_rows, sv.dim_b)
        for d in range(w):
            send[pos[d] : pos[d + 1]].view(sv.n_rows, g["width"][d]).copy_(
                local2d[:, sv.b_off[d] : sv.b_off[d + 1]])
    sv.local = None  # the row shard's storage is dropped before the column shard is allocated
    out_splits = [(sv.a_off[s + 1] - sv.a_off[s]) * sv.n_cols for s in range(w)]
    recv = torch.empty(sv.dim_a * sv.n_cols, dtype=torch.complex128, device=device)
    dist.all_to_all_single(
        torch.view_as_real(recv), torch.view_as_real(send),
        output_split_sizes=out_splits, input_split_sizes=in_splits, group=sv.group,
    )
    sv._adopt(recv, COLS)


def _cols_to_rows_a2a(sv: ShardedVector) -> None:
    """cols -> rows: the send side is contiguous per destination (d's rows of my column shard); the
    receive side is unpacked into the [n_rows x dim_b] shard."""
    w = sv.world
    g = _col_block_geometry(sv)
    device = sv.device
    in_splits = [g["rows"][d] * sv.n_cols for d in range(w)]
    out_width = [sv.b_off[s + 1] - sv.b_off[s] for s in range(w)]
    out_splits = [sv.n_rows * out_width[s] for s in range(w)]
    recv = torch.empty(sv.n_rows * sv.dim_b, dtype=torch.complex128, device=device)
    dist.all_to_all_single(
        torch.view_as_real(recv), torch.view_as_real(sv.local),
        output_split_sizes=out_splits, input_split_sizes=in_splits, group=sv.group,
    )
    sv.local = None
    local = torch.empty(sv.n_rows * sv.dim_b, dtype=torch.complex128, device=device)
    pos = [0]
    for n in out_splits:
        pos.append(pos[-1] + n)
    if recv.is_cuda:
        with torch.cuda.device(device):
            _copy_blocks(recv, out_width, [sv.n_rows] * w, out_width, pos[:-1], [local.data_ptr()] * w,
                         [sv.b_off[s] for s in range(w)], [sv.dim_b] * w)
    else:
        local2d = local.view(sv.n_rows, sv.dim_b)
        for s in range(w):
            local2d[:, sv.b_off[s] : sv.b_off[s + 1]].copy_(recv[pos[s] : pos[s + 1]].view(sv.n_rows, out_width[s]))
    sv._adopt(local, ROWS)


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
_SYMM_STATE = {"checked": False, "ok": False, "sync": None}


def _symm_get(numel: int, device, group) -> _SymmBuffer:
    """Pooled: symmetric allocations are collective and slow, and every rank takes and returns them at
    the same points of the (SPMD) program, so the pools stay identical across ranks."""
    free = _SYMM_POOL.setdefault((id(group), str(device), numel), [])
    return free.pop() if free else _SymmBuffer(numel, device, group)


def _symm_put(buf: _SymmBuffer, device, group) -> None:
    _SYMM_POOL.setdefault((id(group), str(device), buf.numel), []).append(buf)


def _symm_release(box, device, group) -> None:
    """Finalizer of a vector that lives in symmetric memory: its current buffer goes back to the pool."""
    if box[0] is not None:
        _symm_put(box[0], device, group)
        box[0] = None


def p2p_available(sv: "ShardedVector") -> bool:
    """Peer-memory exchange needs CUDA shards, more than one rank, at most 16 of them on one box with
    symmetric memory working; the decision is taken collectively so that all ranks agree."""
    if sv.world == 1 or sv.world > 16 or not sv.device.type == "cuda":
        return False
    mode = os.environ.get("FFSIM_B200_EXCHANGE", "auto").lower()
    # "auto": the configurations validated on this pool -- up to 8 ranks (2, 4, 8 measured) and symmetric
    # buffers up to 16 GB per rank (8 GB measured); larger shards (the 254 GB state: 32 GB) go through NCCL
    max_world = int(os.environ.get("FFSIM_B200_P2P_MAX_WORLD", "8"))
    max_bytes = int(os.environ.get("FFSIM_B200_P2P_MAX_GB", "16")) << 30
    if mode == "nccl" or (mode != "p2p" and (sv.world > max_world or 16 * _symm_numel(sv) > max_bytes)):
        return False
    if not _SYMM_STATE["checked"]:
        ok = 1
        try:
            probe = _SymmBuffer(16, sv.device, sv.group)
            ok = int(len(probe.ptrs) == sv.world and all(probe.ptrs))
            # every cross-rank barrier of the exchange goes through THIS handle: symmetric allocations that
            # share a memory-pool block share a signal pad, and a barrier keeps per-handle state, so
            # alternating barriers on the handles of (small, pooled) data buffers can deadlock
            _SYMM_STATE["sync"] = probe
        except Exception:
            ok = 0
        flag = torch.tensor([ok], device=sv.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=sv.group)
        _SYMM_STATE["ok"] = bool(flag.item())
        _SYMM_STATE["checked"] = True
    return _SYMM_STATE["ok"]


def _symm_numel(sv: ShardedVector) -> int:
    """One size for both distributions (so that the pool hands the same buffers back and forth)."""
    rows_max = max(sv.a_off[r + 1] - sv.a_off[r] for r in range(sv.world))
    cols_max = max(sv.b_off[r + 1] - sv.b_off[r] for r in range(sv.world))
    return max(rows_max * sv.dim_b, sv.dim_a * cols_max, 1)


def _make_symmetric(sv: ShardedVector) -> _SymmBuffer:
    """Move the shard into symmetric memory (once per vector; it stays there)."""
    if sv._symm is not None:  # set by _adopt exactly when ``local`` is a view of that buffer (an EMPTY view has
        return sv._symm       # no data pointer to compare: ranks with a zero-size shard must take this path too)
    buf = _symm_get(_symm_numel(sv), sv.device, sv.group)
    view = buf.tensor[: sv.local.numel()]
    view.copy_(sv.local)
    sv._adopt(view, sv.layout, buf)
    return buf


def _redistribute_p2p(sv: ShardedVector, layout: str) -> None:
    """One kernel: every rank stores its blocks straight into their final place in the peers' shards."""
    src_buf = _make_symmetric(sv)
    dst_buf = _symm_get(_symm_numel(sv), sv.device, sv.group)
    sync = _SYMM_STATE["sync"].handle
    with torch.cuda.device(sv.device):
        sync.barrier(channel=0)  # peers may still be reading this buffer from an earlier exchange
        if layout == COLS:
            g = _row_block_geometry(sv)
            _copy_blocks(sv.local, sv.dim_b, g["rows"], g["width"], g["row_off"], dst_buf.ptrs, g["col_off"], g["col_ld"])
        else:
            g = _col_block_geometry(sv)
            _copy_blocks(sv.local, sv.n_cols, g["rows"], g["width"], g["col_off"], dst_buf.ptrs, g["row_off"], g["row_ld"])
        sync.barrier(channel=0)  # every block has landed before anyone reads its new shard
    del src_buf  # _adopt hands the old buffer back to the pool
    sv._adopt(dst_buf.tensor[: sv.local_numel(layout)], layout, dst_buf)


def redistribute(sv: ShardedVector, layout: str) -> None:
    """Move ``sv`` to the other distribution (collective)."""
    if layout == sv.layout:
        return
    STATS["exchanges"] += 1
    STATS["bytes_sent"] += all_to_all_bytes(sv)
    if sv.world == 1:
        # one rank: both distributions are the whole matrix
        sv.layout = layout
        return
    timed = STATS.get("time") and sv.device.type == "cuda"
    if timed:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    if p2p_available(sv):
        STATS["mode"] = "p2p"
        _redistribute_p2p(sv, layout)
    else:
        STATS["mode"] = "nccl"
        if layout == COLS:
            _rows_to_cols_a2a(sv)
        else:
            _cols_to_rows_a2a(sv)
    if timed:
        ev[1].record()
        STATS["events"].append(ev)


# ---------------------------------------------------------------------- device ops on shards

def _rotate_beta_local(sv: ShardedVector, plan, stream) -> None:
    """Beta-side rotation of the rows-distribution shard: the contiguous index of the local rows."""
    from ffsim_b200 import _lib

    if sv.n_rows == 0:
        return
    # in place when the beta sector fits one window; otherwise through a transposed copy whose
    # transpositions the library folds into the first and the last pass
    ws_ptr = None
    if not _lib.lib.ffb_plan_beta_in_place(plan.handle):
        ws = torch.empty(sv.dim_b * sv.n_rows, dtype=torch.complex128, device=sv.device)
        ws_ptr = ws.data_ptr()
    _lib.check(_lib.lib.ffb_apply_orbital_rotation_beta_block(
        plan.handle, sv.local.data_ptr(), sv.n_rows, sv.dim_b, ws_ptr, stream))


def _rotate_alpha_local(sv: ShardedVector, plan, stream) -> None:
    """Alpha-side rotation of the cols-distribution shard: the row index of a [dim_a x n_cols] matrix."""
    from ffsim_b200 import _lib

    if sv.n_cols > 0:
        _lib.check(_lib.lib.ffb_apply_orbital_rotation_rows(
            plan.handle, 0, sv.local.data_ptr(), sv.n_cols, sv.n_cols, stream))


def rotate(sv: ShardedVector, mat_a, mat_b) -> None:
    """Orbital rotation of a sharded state, in place (both spin sectors).  The spin sector that is local
    in the current distribution goes first (the two commute); the other one follows after ONE
    redistribution, and the state stays in that distribution."""
    from ffsim_b200 import _device
    from ffsim_b200.gates.orbital_rotation import get_plan

    with torch.cuda.device(sv.device):
        plan = get_plan(sv.norb, sv.nelec, mat_a, mat_b)
        stream = _device.stream_ptr()
        if sv.world == 1:
            # one rank holds everything: both sides are local, no need to flip the label
            if mat_b is not None:
                _rotate_beta_local(sv, plan, stream)
            if mat_a is not None:
                _rotate_alpha_local(sv, plan, stream)
            return
        order = ("b", "a") if sv.layout == ROWS else ("a", "b")
        for side in order:
            if side == "b" and mat_b is not None:
                sv.set_layout(ROWS)
                _rotate_beta_local(sv, plan, stream)
            elif side == "a" and mat_a is not None:
                sv.set_layout(COLS)
                _rotate_alpha_local(sv, plan, stream)
