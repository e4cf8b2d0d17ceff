"""CPU-only tests: C ABI surface, bit-exact tables, host decomposition, plan tables."""

import ctypes
import math
import os
import re
import subprocess

import numpy as np
import pytest

from ffsim_b200 import _lib as L
from ffsim_b200 import cistring as fcis
from ffsim_b200.linalg import givens_decomposition
from oracle import cistring, cref, gates, givens, models, rand

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTCHECK = os.path.join(ROOT, "tests", "hostcheck", "libffsim_b200_hostcheck.so")


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ffsim_b200.h")).read()
    declared = set(re.findall(r"\b(ffb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L.lib, name), f"{name} is declared in the header but not exported"
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    assert L.lib.ffb_version() == 100


def test_no_device_is_reported_not_faked():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    assert L.lib.ffb_device_count() == 0
    import ffsim_b200

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ffsim_b200.apply_orbital_rotation(np.ones(1, complex), np.eye(1), 1, (1, 1))


def test_package_does_not_import_oracle():
    out = subprocess.run(
        ["grep", "-rlE", r"^\s*(from|import)\s+oracle", os.path.join(ROOT, "ffsim_b200")],
        capture_output=True, text=True)
    assert out.stdout.strip() == ""


@pytest.mark.parametrize("norb", range(0, 10))
def test_tables_bit_exact(norb):
    for k in range(0, norb + 1):
        s = fcis.make_strings(norb, k)
        assert s.dtype == np.int64 and np.array_equal(s, cistring.make_strings(range(norb), k))
        occ = fcis.gen_occslst(norb, k)
        assert occ.dtype == np.uint64 and np.array_equal(occ, cistring.gen_occslst(range(norb), k))
        assert np.array_equal(fcis.strs2addr(norb, k, s), np.arange(len(s)))
        for i in range(norb):
            if k >= 1:
                assert np.array_equal(fcis.one_subspace_indices(norb, k, (i,)),
                                      cistring.one_subspace_indices(norb, k, (i,)))
            for j in range(norb):
                if i != j:
                    assert np.array_equal(fcis.zero_one_subspace_indices(norb, k, (i, j)),
                                          cistring.zero_one_subspace_indices(norb, k, (i, j)))


def test_tables_large_sector_matches_oracle():
    for norb, k in [(16, 5), (18, 7), (20, 3)]:
        assert np.array_equal(fcis.make_strings(norb, k), cistring.make_strings(range(norb), k))
        assert np.array_equal(fcis.zero_one_subspace_indices(norb, k, (7, 8)),
                              cistring.zero_one_subspace_indices(norb, k, (7, 8)))


def test_tables_errors():
    h = ctypes.c_void_p()
    assert L.lib.ffb_tables_create(3, 4, ctypes.byref(h)) == L.FFB_EINVAL
    assert "nocc" in L.last_error()
    assert L.lib.ffb_tables_create(-1, 0, ctypes.byref(h)) == L.FFB_EINVAL
    with pytest.raises(ValueError):
        fcis.strs2addr(4, 2, [0b0111])


@pytest.mark.parametrize("n", [0, 1, 2, 3, 5, 8, 12, 16, 20, 24])
def test_givens_decomposition_matches_oracle(n):
    u = rand.random_unitary(n, seed=100 + n) if n else np.zeros((0, 0), complex)
    rots, phases = givens_decomposition(u)
    want_rots, want_phases = givens.givens_decomposition(u)
    assert [(r.i, r.j) for r in rots] == [(i, j) for _, _, i, j in want_rots]
    for r, (c, s, _, _) in zip(rots, want_rots):
        assert abs(r.c - c) < 1e-12 and abs(r.s - s) < 1e-12
    np.testing.assert_allclose(phases, want_phases, atol=1e-12)


def test_givens_decomposition_edge_cases():
    rots, phases = givens_decomposition(np.eye(5))  # tests/python/linalg/givens_test.py:122-128
    assert rots == [] and np.allclose(phases, 1)
    with pytest.raises(ValueError):
        givens_decomposition(np.zeros((2, 3)))
    u = np.load(os.path.join(ROOT, "tests", "golden", "orbital_rotation-0.npy"))
    rots, phases = givens_decomposition(u)
    want_rots, _ = givens.givens_decomposition(u)
    assert len(rots) == len(want_rots) < 28  # exact zeros give fewer rotations
    perm = np.eye(4)[[1, 0, 3, 2]]
    rots, phases = givens_decomposition(perm)
    want_rots, want_phases = givens.givens_decomposition(perm)
    assert [(r.c, r.s, r.i, r.j) for r in rots] == [tuple(w) for w in want_rots]


# ---------------------------------------------------------------- plan tables via the host emulator

def _hostcheck():
    if not os.path.exists(HOSTCHECK):
        subprocess.run(["make", "-C", os.path.dirname(HOSTCHECK)], check=True)
    hc = ctypes.CDLL(HOSTCHECK)
    hc.ffb_hostcheck_apply_side.restype = ctypes.c_int
    hc.ffb_hostcheck_apply_side.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    return hc


def _emulate(norb, k, n_cols, smem, min_cols, sub_window, seed, u=None):
    from ffsim_b200.linalg.givens import _decompose_raw

    hc = _hostcheck()
    rng = np.random.default_rng(seed)
    dim = math.comb(norb, k)
    if u is None:
        u = rand.random_unitary(norb, seed=rng)
    vec = rand.random_state_vector(dim * n_cols, seed=rng).reshape(dim, n_cols)
    rots, ph = _decompose_raw(u)
    rots = np.ascontiguousarray(rots)
    got = np.ascontiguousarray(vec.copy())
    n_pass, n_sub = ctypes.c_int(), ctypes.c_int()
    rc = hc.ffb_hostcheck_apply_side(norb, k, L.ptr(rots), len(rots), L.ptr(ph), L.ptr(got), n_cols, smem,
                                     min_cols, sub_window, ctypes.byref(n_pass), ctypes.byref(n_sub))
    assert rc == 0, f"hostcheck invariant {rc} failed for norb={norb} k={k}"
    want = vec.copy()
    gates._rotate_one_spin(want, givens.givens_decomposition(u), norb, k)
    err = np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300)
    return err, n_pass.value, n_sub.value


@pytest.mark.parametrize("norb", range(1, 9))
def test_plan_tables_all_small_sectors(norb):
    for k in range(0, norb + 1):
        for smem, mc, sw in [(220 * 1024, 3, 6), (2048, 4, 3), (1024, 2, 2), (4096, 1, 4), (8192, 4, 5)]:
            err, _, _ = _emulate(norb, k, 3, smem, mc, sw, seed=norb * 100 + k)
            assert err < 1e-13, (norb, k, smem, mc, sw, err)


@pytest.mark.parametrize("norb,k", [(22, 19), (22, 20), (24, 22), (24, 23), (23, 20), (24, 2)])
def test_plan_tables_nearly_filled_large_norb(norb, k):
    """sectors of norb >= 22 with nocc >= norb - 3 used to reach more than kMaxLow
    electrons below a register block (an assert, i.e. an abort of the host process).  The window is now
    capped so that the offset tables always suffice; the emulated kernel must reproduce the oracle."""
    err, n_pass, _ = _emulate(norb, k, 1, 220 * 1024, 3, 6, seed=norb + k)
    assert err < 1e-12 and n_pass >= 1


@pytest.mark.parametrize("norb,k,smem,max_passes", [
    (12, 6, 220 * 1024, 1), (12, 6, 16 * 1024, 4), (14, 5, 32 * 1024, 4), (16, 5, 220 * 1024, 1),
    (16, 8, 220 * 1024, 4)])
def test_plan_tables_baseline_shapes(norb, k, smem, max_passes):
    err, n_pass, n_sub = _emulate(norb, k, 1, smem, 3, 6, seed=7)
    assert err < 1e-13
    assert 1 <= n_pass <= max_passes
    assert n_sub <= norb * (norb - 1) // 2


def test_plan_tables_sparse_unitary():
    u = np.load(os.path.join(ROOT, "tests", "golden", "orbital_rotation-0.npy"))
    err, _, _ = _emulate(8, 5, 2, 4096, 2, 4, seed=3, u=u)
    assert err < 1e-13
    err, _, _ = _emulate(6, 3, 2, 220 * 1024, 3, 6, seed=3, u=np.eye(6))
    assert err < 1e-15


# ---------------------------------------------------------------- the C restatement used as CPU baseline

@pytest.mark.parametrize("norb,nelec", [(5, (3, 2)), (8, (4, 3)), (6, (0, 3)), (4, (4, 2)), (7, (2, 5))])
def test_c_restatement_matches_numpy_oracle(norb, nelec):
    rng = np.random.default_rng(norb)
    v = rand.random_state_vector(models.dim(norb, nelec), seed=rng)
    u = rand.random_unitary(norb, seed=rng)
    np.testing.assert_allclose(cref.apply_orbital_rotation(v, u, norb, nelec),
                               gates.apply_orbital_rotation(v, u, norb, nelec), atol=1e-13)
    mats = (rand.random_real_symmetric_matrix(norb, seed=rng), rng.standard_normal((norb, norb)),
            rand.random_real_symmetric_matrix(norb, seed=rng))
    for z in (False, True):
        np.testing.assert_allclose(
            cref.apply_diag_coulomb_evolution(v, mats, 0.3, norb, nelec, z_representation=z),
            gates.apply_diag_coulomb_evolution(v, mats, 0.3, norb, nelec, z_representation=z), atol=1e-13)


# ----------------------------------------------------------------------------- host logic
def _unitaries(shape, seed):
    rng = np.random.default_rng(seed)
    n = shape[-1]
    out = np.empty(shape, dtype=complex)
    for idx in np.ndindex(*shape[:-2]):
        out[idx] = rand.random_unitary(n, seed=rng)
    return out


def test_ucj_spin_unbalanced_validation_errors():
    import ffsim_b200 as ffsim

    norb, n_reps = 4, 2
    rng = np.random.default_rng(5)
    mats = np.stack([np.stack([rand.random_real_symmetric_matrix(norb, seed=rng), rng.standard_normal((norb, norb)),
                               rand.random_real_symmetric_matrix(norb, seed=rng)]) for _ in range(n_reps)])
    rots = _unitaries((n_reps, 2, norb, norb), 6)
    op = ffsim.UCJOpSpinUnbalanced(mats, rots, final_orbital_rotation=rots[0])  # J_ab need not be symmetric
    assert (op.norb, op.n_reps) == (norb, n_reps)
    with pytest.raises(ValueError, match="shape"):
        ffsim.UCJOpSpinUnbalanced(mats[:, :2], rots)
    with pytest.raises(ValueError, match="shape"):
        ffsim.UCJOpSpinUnbalanced(mats, rots[:, 0])
    with pytest.raises(ValueError, match="shape"):
        ffsim.UCJOpSpinUnbalanced(mats, rots, final_orbital_rotation=rots[0, 0])
    with pytest.raises(ValueError, match="first dimension"):
        ffsim.UCJOpSpinUnbalanced(mats, rots[:1])
    with pytest.raises(ValueError, match="unitary"):
        ffsim.UCJOpSpinUnbalanced(mats, 1.1 * rots)
    with pytest.raises(ValueError, match="unitary"):
        ffsim.UCJOpSpinUnbalanced(mats, rots, final_orbital_rotation=1.1 * rots[0])
    bad = mats.copy()
    bad[1, 2, 0, 1] += 1.0
    with pytest.raises(ValueError, match="symmetric"):
        ffsim.UCJOpSpinUnbalanced(bad, rots)
    ffsim.UCJOpSpinUnbalanced(bad, rots, validate=False)
    # an integer nelec is not supported by this operator: apply_unitary reports it (apply_unitary_protocol.py:76-88)
    with pytest.raises(TypeError):
        ffsim.apply_unitary(np.zeros(6, dtype=complex), op, norb=norb, nelec=2)


def test_ucj_spinless_validation_errors():
    import ffsim_b200 as ffsim

    norb, n_reps = 4, 2
    rng = np.random.default_rng(7)
    mats = np.stack([rand.random_real_symmetric_matrix(norb, seed=rng) for _ in range(n_reps)])
    rots = _unitaries((n_reps, norb, norb), 8)
    op = ffsim.UCJOpSpinless(mats, rots, final_orbital_rotation=rots[1])
    assert (op.norb, op.n_reps) == (norb, n_reps)
    with pytest.raises(ValueError, match="shape"):
        ffsim.UCJOpSpinless(mats[0], rots)
    with pytest.raises(ValueError, match="shape"):
        ffsim.UCJOpSpinless(mats, rots[0])
    with pytest.raises(ValueError, match="shape"):
        ffsim.UCJOpSpinless(mats, rots, final_orbital_rotation=rots)
    with pytest.raises(ValueError, match="first dimension"):
        ffsim.UCJOpSpinless(mats[:1], rots)
    with pytest.raises(ValueError, match="unitary"):
        ffsim.UCJOpSpinless(mats, 0.9 * rots)
    bad = mats.copy()
    bad[0, 0, 1] += 1.0
    with pytest.raises(ValueError, match="symmetric"):
        ffsim.UCJOpSpinless(bad, rots)
    ffsim.UCJOpSpinless(bad, rots, validate=False)


@pytest.mark.parametrize("z_rep", [False, True])
@pytest.mark.parametrize("rank_one", [False, True])
def test_qdrift_probabilities_match_oracle(z_rep, rank_one):
    """Host arithmetic of python/ffsim/trotter/qdrift.py:244-455 ("norm" exact and loose branches, "uniform")."""
    import ffsim_b200 as ffsim

    norb, nelec, rank = 5, (3, 2), 4
    rng = np.random.default_rng(11)
    one_body = rand.random_hermitian(norb, seed=rng) if hasattr(rand, "random_hermitian") else None
    if one_body is None:
        m = rng.standard_normal((norb, norb)) + 1j * rng.standard_normal((norb, norb))
        one_body = m + m.T.conj()
    if rank_one:
        vs = rng.standard_normal((rank, norb))
        mats = np.stack([np.outer(v, v) for v in vs])
    else:
        mats = np.stack([rand.random_real_symmetric_matrix(norb, seed=rng) for _ in range(rank)])
    rots = _unitaries((rank, norb, norb), 12)
    ham = ffsim.DoubleFactorizedHamiltonian(one_body, mats, rots, constant=0.5, z_representation=z_rep)
    for method in ("norm", "uniform"):
        got = ffsim.qdrift_probabilities(ham, method, nelec=nelec)
        want = models.qdrift_probabilities(one_body, mats, z_rep, method, nelec)
        assert got.shape == (rank + 1,) and abs(got.sum() - 1) < 1e-14
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)
    with pytest.raises(ValueError, match="nelec"):
        ffsim.qdrift_probabilities(ham, "norm")
    with pytest.raises(ValueError, match="one_rdm"):
        ffsim.qdrift_probabilities(ham, "optimal")
    with pytest.raises(ValueError, match="Unsupported"):
        ffsim.qdrift_probabilities(ham, "nonsense")


@pytest.mark.parametrize("norb, k, bound", [(16, 5, 0.36), (18, 7, 0.67), (20, 8, 0.52), (24, 6, 0.55)])
def test_block_lists_are_ordered_for_conflict_free_gathers(norb, k, bound):
    """plan.cpp orders the blocks of a class so that, within one l' group, eight consecutive blocks
    start in different bank groups.  What is left are the group tails (the residues of a group are not
    evenly populated); the bounds are the values of round 1 (+0.05) and guard against regressions of the
    ordering -- the natural (H', r) order gives 0.72, 0.90, 0.73, 0.72 extra wavefronts per quarter-warp."""
    hc = _hostcheck()
    hc.ffb_hostcheck_gather_conflicts.restype = ctypes.c_int
    hc.ffb_hostcheck_gather_conflicts.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                                  ctypes.POINTER(ctypes.c_int64)]
    rots, _ = givens_decomposition(rand.random_unitary(norb, seed=3))
    q = (ctypes.c_int * len(rots))(*[min(int(r[2]), int(r[3])) for r in rots])
    out = (ctypes.c_int64 * 2)()
    assert hc.ffb_hostcheck_gather_conflicts(norb, k, q, len(rots), out) == 0
    quarters, extra = out[0], out[1]
    assert quarters > 100
    assert extra <= bound * quarters, f"{extra} extra wavefronts in {quarters} quarter-warps"


@pytest.mark.parametrize("norb, k, cols, bound", [(16, 5, 3, 1.25), (18, 7, 0, 1.20), (20, 8, 0, 1.22), (12, 6, 8, 1.001)])
def test_gather_wavefronts_with_column_fastest_items(norb, k, cols, bound):
    """The register-block gathers and scatters as the kernel walks them (32-item chunks, a quarter-warp per
    shared-memory wavefront): with the items of a chunk ordered column-fastest (device_structs.h:
    items_column_fastest) the wavefronts per conflict-free wavefront are 1.20 (C2), 1.14 (C3), 1.17 (C4), 1.00 (C1,
    8 columns) -- and at least a tenth below the block-fastest walk of the same lists (1.40, 1.73, 1.54, 1.92)."""
    hc = _hostcheck()
    hc.ffb_hostcheck_gather_wavefronts.restype = ctypes.c_int
    hc.ffb_hostcheck_gather_wavefronts.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
    rots, _ = givens_decomposition(rand.random_unitary(norb, seed=1))
    q = (ctypes.c_int * len(rots))(*[min(int(r[2]), int(r[3])) for r in rots])
    ratio = {}
    for col_fast in (0, 1):
        out = (ctypes.c_int64 * 2)()
        assert hc.ffb_hostcheck_gather_wavefronts(norb, k, q, len(rots), cols, col_fast, out) == 0
        assert out[1] > 1000
        ratio[col_fast] = out[0] / out[1]
    assert ratio[1] <= bound, ratio
    assert ratio[1] <= 0.9 * ratio[0], ratio


# ---------------------------------------------------------------- the kernel's index path, from the packed device tables
def _emulate_device_view(norb, k, n_cols, smem, min_cols, sub_window, cols, nwarp, seed):
    from ffsim_b200.linalg.givens import _decompose_raw

    hc = _hostcheck()
    hc.ffb_hostcheck_apply_side_device.restype = ctypes.c_int
    hc.ffb_hostcheck_apply_side_device.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    rng = np.random.default_rng(seed)
    dim = math.comb(norb, k)
    u = rand.random_unitary(norb, seed=rng)
    vec = rand.random_state_vector(dim * n_cols, seed=rng).reshape(dim, n_cols)
    rots, _ = _decompose_raw(u)
    rots = np.ascontiguousarray(rots)
    got = np.ascontiguousarray(vec.copy())
    rc = hc.ffb_hostcheck_apply_side_device(norb, k, L.ptr(rots), len(rots), L.ptr(got), n_cols, smem, min_cols,
                                            sub_window, cols, nwarp)
    assert rc == 0, f"device-view invariant {rc} failed for norb={norb} k={k}"
    want = vec.copy()
    rotations, _ = givens.givens_decomposition(u)
    gates._rotate_one_spin(want, (rotations, []), norb, k)  # the phases are a separate kernel stage
    return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300)


@pytest.mark.parametrize("norb,k,n_cols,smem,min_cols,sub_window,cols,nwarp", [
    (6, 3, 5, 220 * 1024, 3, 6, 3, 16),      # one window, ragged last column strip
    (8, 4, 7, 220 * 1024, 3, 6, 8, 16),      # eight tile columns (stride rule for cols = 8)
    (8, 3, 4, 4096, 2, 4, 2, 4),             # several passes and tile groups, 4-orbital blocks, 4 warps
    (9, 4, 3, 8192, 1, 5, 1, 16),            # single-column tiles
    (10, 5, 6, 220 * 1024, 3, 6, 3, 16),
    (10, 2, 9, 2048, 4, 3, 4, 8),
    (12, 6, 2, 220 * 1024, 3, 6, 2, 16),     # BASELINE C1 sector
    (14, 5, 3, 32 * 1024, 3, 6, 3, 16),
    (16, 5, 1, 220 * 1024, 3, 6, 1, 16),     # BASELINE C2 sector, one column
    (7, 7, 2, 220 * 1024, 3, 6, 2, 16),      # full shell: nothing to rotate
    (5, 1, 3, 220 * 1024, 3, 6, 3, 12),
])
def test_device_view_of_the_plan_tables(norb, k, n_cols, smem, min_cols, sub_window, cols, nwarp):
    """tests/hostcheck/emulate.cpp::ffb_hostcheck_apply_side_device walks the packed device tables the way
    fused_pass_kernel does (chunk dealing, multiply-high division, byte offsets, column-major tile) and must
    reproduce the oracle's rotation."""
    err = _emulate_device_view(norb, k, n_cols, smem, min_cols, sub_window, cols, nwarp, seed=100 * norb + k)
    assert err < 1e-13


# ---------------------------------------------------------------- bench.py: the reference (CPU) arm's contract line
def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` runs here (no GPU): exactly one line on stdout, a JSON object with the
    contract's keys, the CPU baseline described, zero copy bytes; ranks other than 0 print nothing."""
    import json
    import subprocess
    import sys

    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c2", "--steps", "1",
           "--warmup", "0"]
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    rec = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in rec, key
    assert rec["impl"] == "reference" and rec["value"] > 0 and rec["unit"] == "applications/s"
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["d2h_bytes_per_step"] == 0
    assert "norb=16" in rec["config"]["workload"]
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert other.returncode == 0 and other.stdout.strip() == ""
