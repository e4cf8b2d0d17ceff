"""Host-resident states: a sequence of hot-path operations with the PCIe copies overlapped with the kernels.

``evolve_host(vec, steps, norb, nelec)`` is equivalent to calling the public functions one after the other
on a NumPy array -- but the state crosses PCIe once in each direction and most of the kernel time hides
behind the copies:

* an orbital rotation acts on every beta column independently on its alpha side and on every alpha row
  independently on its beta side; a diagonal operator acts on every amplitude independently;
* so the state is **uploaded in column strips**, and the leading column-local operations (the alpha side of
  the first rotation, leading diagonal operators) run on each strip as soon as it has arrived, while the
  next strip is in flight;
* operations in the middle run on the whole device-resident state;
* the trailing row-local operations (the beta side of the last rotation, trailing diagonal operators) run
  **row block by row block**, and each finished block is downloaded while the next one is computed.

The reference has no counterpart (its states live in host memory and so do its kernels); this is the
end-to-end path of a B200 drop-in: a 16.2 GB state spends 2 x 300 ms on PCIe and 220 ms in kernels.
Copies use ``ffb_memcpy2d_async`` (cudaMemcpy2DAsync); page-locked host memory is needed for them to be
asynchronous (``ffsim_b200.pinned_empty``), pageable memory works but serialises.
"""

from __future__ import annotations

import numpy as np
import torch

from ffsim_b200 import _device, _lib
from ffsim_b200.cistring import get_tables
from ffsim_b200.distributed import partition
from ffsim_b200.gates.diag_coulomb import _get_mat_exp
from ffsim_b200.gates.num_op_sum import _get_phases
from ffsim_b200.gates.orbital_rotation import _split_mat, get_plan

_H2D, _D2H = 1, 2


def pinned_empty(n: int) -> np.ndarray:
    """A page-locked ``complex128`` NumPy array of ``n`` entries (from the package's pinned pool)."""
    buf = _device._pool_get(16 * int(n))
    return np.asarray(_device._PinnedOwner(buf, int(n)))


class _Prim:
    """One primitive: the alpha side or the beta side of a rotation, or a diagonal operator."""

    def __init__(self, kind, **kw):
        self.kind = kind  # "alpha" | "beta" | "diag_coulomb" | "num_op_sum"
        self.__dict__.update(kw)

    @property
    def column_local(self) -> bool:
        return self.kind != "beta"

    @property
    def row_local(self) -> bool:
        return self.kind != "alpha"


def _expand(steps, norb: int, nelec) -> list[_Prim]:
    prims: list[_Prim] = []
    for step in steps:
        name = step[0]
        if name == "orbital_rotation":
            mats = _split_mat(step[1])
            # (plans with the same structure share one cached object whose coefficients are re-installed by
            # get_plan: a primitive therefore carries its matrices, and the runner installs them on use)
            if mats[0] is not None:
                prims.append(_Prim("alpha", mats=mats, rotation=len(prims)))
            if mats[1] is not None:
                prims.append(_Prim("beta", mats=mats, rotation=len(prims) - (mats[0] is not None)))
        elif name == "diag_coulomb":
            z = bool(step[3]) if len(step) > 3 else False
            prims.append(_Prim("diag_coulomb", mats=_get_mat_exp(step[1], float(step[2]), norb, z), z=z))
        elif name == "num_op_sum":
            pa, pb = _get_phases(step[1], float(step[2]))
            prims.append(_Prim("num_op_sum", phases=(pa, pb)))
        else:
            raise ValueError(f"unknown step {name!r}: expected 'orbital_rotation', 'diag_coulomb' or 'num_op_sum'")
    return prims


class _Runner:
    """Applies primitives to a rectangular block of the device-resident (dim_a x dim_b) state."""

    def __init__(self, dev: torch.Tensor, norb: int, nelec, stream: int):
        self.dev, self.norb, self.nelec, self.stream = dev, norb, nelec, stream
        self.ta, self.tb = get_tables(norb, nelec[0]), get_tables(norb, nelec[1])
        self.dim_a, self.dim_b = self.ta.dim, self.tb.dim
        self.base = dev.data_ptr()
        self._installed, self._plan = None, None

    def plan_for(self, prim: _Prim):
        if self._installed != prim.rotation:
            self._plan = get_plan(self.norb, self.nelec, prim.mats[0], prim.mats[1])
            self._installed = prim.rotation
        return self._plan

    def ptr(self, row: int, col: int) -> int:
        return self.base + 16 * (row * self.dim_b + col)

    def run(self, prim: _Prim, r0: int, r1: int, c0: int, c1: int) -> None:
        """``prim`` on rows [r0, r1) x columns [c0, c1); an alpha prim needs all rows, a beta prim all columns."""
        L, st = _lib.lib, self.stream
        n_rows, n_cols = r1 - r0, c1 - c0
        if n_rows <= 0 or n_cols <= 0:
            return
        if prim.kind == "alpha":
            assert (r0, r1) == (0, self.dim_a)
            plan = self.plan_for(prim)
            _lib.check(L.ffb_apply_orbital_rotation_rows(plan.handle, 0, self.ptr(0, c0), n_cols, self.dim_b, st))
        elif prim.kind == "beta":
            assert (c0, c1) == (0, self.dim_b)
            plan = self.plan_for(prim)
            ws_ptr = None
            if not L.ffb_plan_beta_in_place(plan.handle):  # multi-pass beta side: transposed copy [dim_b x n_rows]
                ws = torch.empty(n_rows * self.dim_b, dtype=torch.complex128, device=self.dev.device)
                ws_ptr = ws.data_ptr()
            _lib.check(L.ffb_apply_orbital_rotation_beta_block(plan.handle, self.ptr(r0, 0), n_rows, self.dim_b,
                                                               ws_ptr, st))
        elif prim.kind == "diag_coulomb":
            aa, ab, bb = prim.mats
            _lib.check(L.ffb_apply_diag_coulomb_evolution_block(
                self.ta.handle, self.tb.handle, _lib.ptr(aa), _lib.ptr(ab), _lib.ptr(bb), int(prim.z),
                self.ptr(r0, c0), r0, n_rows, c0, n_cols, self.dim_b, st))
        else:
            pa, pb = prim.phases
            _lib.check(L.ffb_apply_num_op_sum_evolution_block(
                self.ta.handle, self.tb.handle, _lib.ptr(pa), _lib.ptr(pb), self.ptr(r0, c0), r0, n_rows, c0, n_cols,
                self.dim_b, st))


def _copy2d(dst: int, dst_pitch: int, src: int, src_pitch: int, width: int, height: int, kind: int, stream: int) -> None:
    _lib.check(_lib.lib.ffb_memcpy2d_async(dst, dst_pitch, src, src_pitch, width, height, kind, stream))


class _Lane:
    """Per-device resources of the host pipeline: the two copy streams (one per PCIe direction) and a ring of
    device buffers, so that consecutive applications overlap -- the upload of the next state runs while the
    previous one is still being computed and downloaded."""

    def __init__(self, dev_index: int):
        self.dev_index = dev_index
        self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
        self.ring: dict[int, list] = {}  # n_complex -> [[tensor, event recorded after the last copy out of it], ...]
        self.turn: dict[int, int] = {}

    def acquire(self, n: int, depth: int):
        ring = self.ring.setdefault(n, [])
        if len(ring) < depth:
            ring.append([torch.empty(n, dtype=torch.complex128, device="cuda"), None])
            slot = ring[-1]
        else:
            k = self.turn.get(n, 0) % len(ring)
            self.turn[n] = k + 1
            slot = ring[k]
        return slot

    def trim(self, keep_n: int) -> None:
        for n in [n for n in self.ring if n != keep_n]:
            del self.ring[n]


_LANES: dict[int, _Lane] = {}
_DEPTH = 2  # device buffers per state size: one being filled while the other is computed on / drained


class HostEvolution:
    """Handle of an application started by :func:`evolve_host_async`; ``result()`` waits for the download."""

    def __init__(self, out: np.ndarray, done: "torch.cuda.Event | None", keep):
        self._out, self._done, self._keep = out, done, keep

    def done(self) -> bool:
        return self._done is None or self._done.query()

    def result(self) -> np.ndarray:
        if self._done is not None:
            self._done.synchronize()
            self._done, self._keep = None, None
        return self._out


def evolve_host_async(vec, steps, norb: int, nelec: tuple[int, int], *, n_chunks: int | None = None) -> HostEvolution:
    """Start applying ``steps`` to the host vector ``vec``; returns at once with a :class:`HostEvolution`.

    ``steps`` is a list of tuples, applied in order:

    * ``("orbital_rotation", mat)``                         -- as ``apply_orbital_rotation(vec, mat, ...)``
    * ``("diag_coulomb", mat, time[, z_representation])``   -- as ``apply_diag_coulomb_evolution``
    * ``("num_op_sum", coeffs, time)``                      -- as ``apply_num_op_sum_evolution``

    The input is not modified, but it must stay untouched until ``result()`` has returned (it is read by
    asynchronous copies; pageable memory makes them synchronous).  ``n_chunks`` (default: one per ~64 MB, at
    most 16) is the number of column strips / row blocks the copies are cut into.  Applications started back
    to back overlap: uploads, kernels and downloads each run in order on their own stream, and the device
    state buffers form a ring of two, so the upload of one application proceeds while the previous one is
    computed and sent back (both PCIe directions busy at once).
    """
    _device.require_cuda()
    nelec = (int(nelec[0]), int(nelec[1]))
    host = np.ascontiguousarray(np.asarray(vec).reshape(-1), dtype=np.complex128)
    ta, tb = get_tables(norb, nelec[0]), get_tables(norb, nelec[1])
    dim_a, dim_b = ta.dim, tb.dim
    if host.size != dim_a * dim_b:
        raise ValueError(f"vec has {host.size} entries, expected {dim_a * dim_b} for norb={norb}, nelec={nelec}")
    if host.size == 0:
        return HostEvolution(host.copy(), None, None)
    dev_index = _device.sync_device()
    prims = _expand(steps, norb, nelec)
    nbytes = 16 * host.size
    if n_chunks is None:
        n_chunks = max(1, min(16, nbytes // (64 << 20)))
    n_chunks = max(1, min(int(n_chunks), dim_a, dim_b))

    # head: leading column-local primitives; tail: trailing row-local ones; the rest runs on the whole state
    n_head = 0
    while n_head < len(prims) and prims[n_head].column_local:
        n_head += 1
    n_tail = 0
    while n_tail < len(prims) - n_head and prims[len(prims) - 1 - n_tail].row_local:
        n_tail += 1
    head, middle, tail = prims[:n_head], prims[n_head : len(prims) - n_tail], prims[len(prims) - n_tail :]

    with torch.cuda.device(dev_index):
        lane = _LANES.get(dev_index)
        if lane is None:
            lane = _LANES[dev_index] = _Lane(dev_index)
        lane.trim(host.size)  # one state size at a time keeps its ring; other sizes give their memory back
        main = torch.cuda.current_stream()
        s_in, s_out = lane.s_in, lane.s_out
        slot = lane.acquire(host.size, _DEPTH)
        dev = slot[0]
        dev.record_stream(s_in)  # the caching allocator must not hand the buffer on while copies are pending
        dev.record_stream(s_out)
        if slot[1] is not None:
            s_in.wait_event(slot[1])  # the previous user of this buffer has been copied out
        else:
            s_in.wait_stream(main)  # a fresh allocation is ordered on the current stream
        out = pinned_empty(host.size)
        run = _Runner(dev, norb, nelec, main.cuda_stream)
        src, dst, pitch = host.ctypes.data, out.ctypes.data, 16 * dim_b
        everything_in_head = not middle and not tail

        # ---- upload in column strips; head primitives per strip
        col_off = partition(dim_b, n_chunks)
        for k in range(n_chunks):
            c0, c1 = col_off[k], col_off[k + 1]
            if c1 == c0:
                continue
            _copy2d(run.ptr(0, c0), pitch, src + 16 * c0, pitch, 16 * (c1 - c0), dim_a, _H2D, s_in.cuda_stream)
            arrived = torch.cuda.Event()
            arrived.record(s_in)
            main.wait_event(arrived)
            for prim in head:
                run.run(prim, 0, dim_a, c0, c1)
            if everything_in_head:  # nothing couples the strips any more: send each one straight back
                done = torch.cuda.Event()
                done.record(main)
                s_out.wait_event(done)
                _copy2d(dst + 16 * c0, pitch, run.ptr(0, c0), pitch, 16 * (c1 - c0), dim_a, _D2H, s_out.cuda_stream)

        if not everything_in_head:
            for prim in middle:
                run.run(prim, 0, dim_a, 0, dim_b)
            # ---- tail primitives per row block; download each finished block
            row_off = partition(dim_a, n_chunks)
            for k in range(n_chunks):
                r0, r1 = row_off[k], row_off[k + 1]
                if r1 == r0:
                    continue
                for prim in tail:
                    run.run(prim, r0, r1, 0, dim_b)
                done = torch.cuda.Event()
                done.record(main)
                s_out.wait_event(done)
                _copy2d(dst + 16 * r0 * dim_b, pitch, run.ptr(r0, 0), pitch, 16 * (r1 - r0) * dim_b, 1, _D2H,
                        s_out.cuda_stream)
        finished = torch.cuda.Event()
        finished.record(s_out)
        slot[1] = finished
    return HostEvolution(out, finished, (host, dev))


def evolve_host(vec, steps, norb: int, nelec: tuple[int, int], *, n_chunks: int | None = None) -> np.ndarray:
    """Apply ``steps`` (see :func:`evolve_host_async`) to the host vector ``vec`` and return the result as a
    new (page-locked) NumPy array.  The input is not modified."""
    return evolve_host_async(vec, steps, norb, nelec, n_chunks=n_chunks).result()


def evolve_host_many(vecs, steps, norb: int, nelec: tuple[int, int], *, n_chunks: int | None = None) -> list[np.ndarray]:
    """``[evolve_host(v, steps, ...) for v in vecs]`` with consecutive applications overlapped: while one
    state is computed and downloaded, the next one is already being uploaded (see :func:`evolve_host_async`).
    At most two applications are in flight (two device state buffers); results come back in order."""
    results: list[np.ndarray] = []
    pending: list[HostEvolution] = []
    for v in vecs:
        if len(pending) >= _DEPTH:  # its device buffer is the one the next application needs
            results.append(pending.pop(0).result())
        pending.append(evolve_host_async(v, steps, norb, nelec, n_chunks=n_chunks))
    results.extend(h.result() for h in pending)
    return results


def evolve_host_rows_async(rows, steps, norb: int, nelec: tuple[int, int], *, group=None) -> HostEvolution:
    """The sharded counterpart of :func:`evolve_host_async` (one process per GPU, ``torch.distributed``
    initialised): ``rows`` is THIS rank's block of alpha rows of the host state (``ShardedVector`` row
    distribution, ``(n_rows, dim_b)`` complex128, page-locked for the copies to be asynchronous); the handle's
    ``result()`` is the same block of the evolved state.  Collective: every rank calls it with the same
    ``steps``.  The shard is uploaded on its own stream, the operations run on the current stream through the
    public functions on a ``ShardedVector`` (exchanges over NVLink included), and the result is downloaded on a
    third stream -- so with two applications in flight the upload of one overlaps the kernels and the
    download of the one before, and both PCIe directions stay busy."""
    from ffsim_b200 import apply_diag_coulomb_evolution, apply_num_op_sum_evolution, apply_orbital_rotation
    from ffsim_b200.distributed import ROWS, ShardedVector

    _device.require_cuda()
    nelec = (int(nelec[0]), int(nelec[1]))
    host = np.ascontiguousarray(np.asarray(rows).reshape(-1), dtype=np.complex128)
    dev_index = _device.sync_device()
    with torch.cuda.device(dev_index):
        lane = _LANES.get(dev_index)
        if lane is None:
            lane = _LANES[dev_index] = _Lane(dev_index)
        main = torch.cuda.current_stream()
        local = torch.empty(host.size, dtype=torch.complex128, device="cuda")
        sv = ShardedVector(local, norb, nelec, group)  # checks the shard size
        out = pinned_empty(host.size)
        if host.size:
            lane.s_in.wait_stream(main)  # the allocation is ordered on the current stream
            local.record_stream(lane.s_in)
            _copy2d(local.data_ptr(), 16 * host.size, host.ctypes.data, 16 * host.size, 16 * host.size, 1, _H2D,
                    lane.s_in.cuda_stream)
            arrived = torch.cuda.Event()
            arrived.record(lane.s_in)
            main.wait_event(arrived)
        for step in steps:
            name = step[0]
            if name == "orbital_rotation":
                apply_orbital_rotation(sv, step[1], norb, nelec, copy=False)
            elif name == "diag_coulomb":
                apply_diag_coulomb_evolution(sv, step[1], float(step[2]), norb, nelec,
                                             z_representation=bool(step[3]) if len(step) > 3 else False, copy=False)
            elif name == "num_op_sum":
                apply_num_op_sum_evolution(sv, step[1], float(step[2]), norb, nelec, copy=False)
            else:
                raise ValueError(f"unknown step {name!r}: expected 'orbital_rotation', 'diag_coulomb' or 'num_op_sum'")
        sv.set_layout(ROWS)  # the caller's buffer holds rows: same distribution out as in
        finished = None
        if host.size:
            done = torch.cuda.Event()
            done.record(main)
            lane.s_out.wait_event(done)
            sv.local.record_stream(lane.s_out)
            _copy2d(out.ctypes.data, 16 * host.size, sv.local.data_ptr(), 16 * host.size, 16 * host.size, 1, _D2H,
                    lane.s_out.cuda_stream)
            finished = torch.cuda.Event()
            finished.record(lane.s_out)
    # the vector (and with it its symmetric-memory buffers) stays out of the pools until the result is collected
    return HostEvolution(out, finished, (host, sv))


def release_device_buffers() -> None:
    """Free the device state buffers the host pipeline keeps between calls."""
    torch.cuda.synchronize()
    for lane in _LANES.values():
        lane.ring.clear()
        lane.turn.clear()
    torch.cuda.empty_cache()
