"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerance: relative 2-norm error <= 1e-12 in fp64 (BASELINE.json north_star).
"""

import math
import os

import numpy as np
import pytest
import scipy.sparse.linalg

pytestmark = pytest.mark.gpu

import ffsim_b200 as ffsim  # noqa: E402
from ffsim_b200 import _lib  # noqa: E402
from ffsim_b200.gates.orbital_rotation import (  # noqa: E402
    apply_givens_rotation_in_place, apply_orbital_rotation_unfused, apply_phase_shift_in_place, get_plan)
from oracle import cistring, compound, contract, cref, gates, models, rand  # noqa: E402

TOL = 1e-12
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

NORB_NELEC_CASES = [  # python/ffsim/testing/testing.py:24-35
    (0, (0, 0)), (1, (0, 0)), (1, (0, 1)), (1, (1, 0)), (1, (1, 1)),
    (2, (0, 0)), (2, (2, 2)), (3, (1, 2)), (4, (2, 2)), (4, (3, 2)),
]
MEDIUM_CASES = [(5, (3, 2)), (6, (3, 3)), (7, (2, 5)), (8, (4, 4)), (9, (3, 4)), (10, (5, 5)), (11, (2, 3))]
NORB_NOCC_CASES = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2), (3, 1), (4, 2), (5, 3)]  # testing.py:37-46


def rel_err(got, want):
    n = np.linalg.norm(want)
    return np.linalg.norm(np.asarray(got) - want) / (n if n > 0 else 1.0)


@pytest.fixture(autouse=True)
def _default_options():
    saved = {k: _lib.get_option(k) for k in ("smem_bytes", "min_cols", "max_cols", "sub_window", "threads", "beta_mode", "bulk_copies")}
    yield
    for k, v in saved.items():
        _lib.set_option(k, v)


def _state(norb, nelec, rng):
    return rand.random_state_vector(max(models.dim(norb, nelec), 1), seed=rng)


# ------------------------------------------------------------------ _lib-level kernels

def test_single_givens_and_phase_shift_kernels():
    # tests/python/_slow/gates/orbital_rotation_test.py:27-46 (norb=5, general target orbitals)
    import torch

    rng = np.random.default_rng(1)
    norb, nelec = 5, (3, 2)
    dim_a, dim_b = models.dims(norb, nelec)
    vec = rand.random_state_vector(dim_a * dim_b, seed=rng).reshape(dim_a, dim_b)
    c = 0.6
    s = 0.8 * np.exp(0.3j)
    idx = cistring.zero_one_subspace_indices(norb, nelec[0], (1, 3))
    half = len(idx) // 2
    want = vec.copy()
    gates.apply_givens_rotation_in_place(want, c, s, idx[:half], idx[half:])
    t = torch.from_numpy(vec.copy()).cuda()
    apply_givens_rotation_in_place(t, c, s, idx[:half], idx[half:])
    assert rel_err(t.cpu().numpy(), want) < TOL
    one = cistring.one_subspace_indices(norb, nelec[0], (2,))
    gates.apply_phase_shift_in_place(want, np.exp(0.7j), one)
    apply_phase_shift_in_place(t, np.exp(0.7j), one)
    assert rel_err(t.cpu().numpy(), want) < TOL
    apply_givens_rotation_in_place(t, c, s, idx[:0], idx[:0])  # empty slice is a no-op (orbital_rotation.rs:27)
    assert rel_err(t.cpu().numpy(), want) < TOL


@pytest.mark.parametrize("norb,nelec", [(4, (2, 2)), (6, (3, 2)), (8, (4, 4))])
def test_unfused_path_matches_oracle(norb, nelec):
    rng = np.random.default_rng(2)
    vec = _state(norb, nelec, rng)
    ua, ub = rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)
    got = apply_orbital_rotation_unfused(vec, (ua, ub), norb, nelec)
    assert rel_err(got, gates.apply_orbital_rotation(vec, (ua, ub), norb, nelec)) < TOL


# ------------------------------------------------------------------ fused orbital rotation

@pytest.mark.parametrize("norb,nelec", NORB_NELEC_CASES + MEDIUM_CASES)
def test_orbital_rotation_spinful(norb, nelec):
    rng = np.random.default_rng(norb * 10 + sum(nelec))
    vec = _state(norb, nelec, rng)
    ua, ub = rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)
    original = vec.copy()
    for mat in (ua, (ua, ub), (ua, None), (None, ub), np.stack([ua, ub])):
        got = ffsim.apply_orbital_rotation(vec, mat, norb, nelec)
        want = gates.apply_orbital_rotation(vec, mat if not isinstance(mat, np.ndarray) or mat.ndim == 2 else tuple(mat),
                                            norb, nelec)
        assert got.shape == want.shape and got.dtype == np.complex128
        assert rel_err(got, want) < TOL
    np.testing.assert_array_equal(vec, original)  # orbital_rotation_test.py:126-137
    np.testing.assert_array_equal(ua, ua.copy())


@pytest.mark.parametrize("norb,nocc", NORB_NOCC_CASES)
def test_orbital_rotation_spinless(norb, nocc):
    rng = np.random.default_rng(norb * 10 + nocc)
    vec = _state(norb, nocc, rng)
    u = rand.random_unitary(norb, seed=rng)
    got = ffsim.apply_orbital_rotation(vec, u, norb, nocc)
    assert rel_err(got, gates.apply_orbital_rotation(vec, u, norb, nocc)) < TOL


@pytest.mark.parametrize("opts", [
    dict(smem_bytes=4096, min_cols=2, sub_window=3),
    dict(smem_bytes=8192, min_cols=4, sub_window=4, beta_mode=2),
    dict(smem_bytes=8192, min_cols=4, sub_window=4, beta_mode=3),
    dict(smem_bytes=4096, min_cols=2, sub_window=5, beta_mode=2),
    dict(smem_bytes=16384, min_cols=1, sub_window=5, beta_mode=1),
    dict(smem_bytes=32768, min_cols=4, sub_window=6, threads=256),
    dict(sub_window=2, threads=128),
    dict(beta_mode=2),
    dict(beta_mode=1, max_cols=4),
])
@pytest.mark.parametrize("norb,nelec", [(6, (3, 2)), (9, (4, 5)), (10, (5, 3)), (12, (3, 3))])
def test_orbital_rotation_plan_variants(opts, norb, nelec):
    """Multi-pass windows, every register-block width, both beta layouts."""
    for k, v in opts.items():
        _lib.set_option(k, v)
    rng = np.random.default_rng(5)
    vec = _state(norb, nelec, rng)
    ua, ub = rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)
    got = ffsim.apply_orbital_rotation(vec, (ua, ub), norb, nelec)
    assert rel_err(got, cref.apply_orbital_rotation(vec, (ua, ub), norb, nelec)) < TOL


def test_beta_transposes_fold_into_the_passes():
    """A multi-pass beta side whose first and last windows start at orbital 0 needs no transpose kernel:
    two state sweeps fewer than with separate transposes (beta_mode=3), same result."""
    from ffsim_b200.gates.orbital_rotation import get_plan

    norb, nelec = 12, (5, 6)
    rng = np.random.default_rng(77)
    vec = _state(norb, nelec, rng)
    u = rand.random_unitary(norb, seed=rng)
    want = cref.apply_orbital_rotation(vec, (None, u), norb, nelec)
    sweeps = {}
    for mode in (2, 3):
        for k, v in dict(smem_bytes=16384, min_cols=2, beta_mode=mode).items():
            _lib.set_option(k, v)
        plan = get_plan(norb, nelec, None, u)
        assert "layout=transposed" in plan.describe() and plan.describe().count("[lo=") >= 2
        sweeps[mode] = plan.n_state_passes()
        assert rel_err(ffsim.apply_orbital_rotation(vec, (None, u), norb, nelec), want) < TOL
    first_lo0 = "transposed [lo=0 " in plan.describe()
    assert sweeps[3] - sweeps[2] >= (1 if first_lo0 else 0)
    assert sweeps[3] - sweeps[2] <= 2


def test_docs_golden_vectors():
    # docs/explanations/state-vectors-and-gates.ipynb cells 9, 13
    u = ffsim.random.random_unitary(3, seed=1234)
    got = ffsim.apply_orbital_rotation(ffsim.hartree_fock_state(3, (2, 1)), u, norb=3, nelec=(2, 1))
    want = np.array([
        0.23611476 + 0.03101213j, -0.06273307 + 0.1102529j, 0.09723851 + 0.36730125j,
        0.13113848 + 0.17276745j, -0.11157654 + 0.02998708j, -0.17558331 + 0.29821173j,
        -0.20881506 - 0.33731417j, 0.20835741 - 0.03525116j, 0.3714141 - 0.51253171j])
    np.testing.assert_allclose(got, want, atol=1e-8)
    got = ffsim.apply_orbital_rotation(ffsim.hartree_fock_state(3, 2), u, norb=3, nelec=2)
    np.testing.assert_allclose(
        got, [-0.4390672 - 0.1561685j, -0.18007105 - 0.38435478j, 0.26121865 + 0.73105542j], atol=1e-8)


def test_orbital_rotation_special_unitaries():
    norb, nelec = 8, (5, 5)
    u = np.load(os.path.join(GOLDEN, "orbital_rotation-0.npy"))  # orbital_rotation_test.py:187-206
    vec = ffsim.hartree_fock_state(norb, nelec)
    got = ffsim.apply_orbital_rotation(vec, u, norb, nelec)
    assert abs(np.linalg.norm(got) - 1) < 1e-12
    minors = compound.slater_minors(u, norb, 5)
    assert rel_err(got, np.outer(minors, minors).reshape(-1)) < TOL
    rng = np.random.default_rng(0)
    vec = _state(6, (3, 3), rng)
    assert rel_err(ffsim.apply_orbital_rotation(vec, np.eye(6), 6, (3, 3)), vec) < 1e-15
    d = np.diag(np.exp(1j * rng.uniform(0, 6, 6)))
    assert rel_err(ffsim.apply_orbital_rotation(vec, d, 6, (3, 3)),
                   gates.apply_orbital_rotation(vec, d, 6, (3, 3))) < TOL
    perm = np.eye(6)[[1, 0, 2, 4, 3, 5]].astype(complex)
    assert rel_err(ffsim.apply_orbital_rotation(vec, perm, 6, (3, 3)),
                   gates.apply_orbital_rotation(vec, perm, 6, (3, 3))) < TOL


def test_orbital_rotation_composition_and_compound_oracle():
    # orbital_rotation_test.py:165-184 and the Givens-free closed form
    norb, nelec = 7, (3, 4)
    rng = np.random.default_rng(8)
    vec = _state(norb, nelec, rng)
    u1, u2 = rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)
    a = ffsim.apply_orbital_rotation(vec, u1, norb, nelec)
    a = ffsim.apply_orbital_rotation(a, u2 @ u1.T.conj(), norb, nelec)
    b = ffsim.apply_orbital_rotation(vec, u2, norb, nelec)
    assert rel_err(a, b) < TOL
    assert rel_err(b, compound.apply_orbital_rotation_compound(vec, u2, norb, nelec)) < TOL


def test_device_tensor_in_device_tensor_out():
    import torch

    norb, nelec = 6, (3, 2)
    rng = np.random.default_rng(4)
    vec = _state(norb, nelec, rng)
    u = rand.random_unitary(norb, seed=rng)
    want = gates.apply_orbital_rotation(vec, u, norb, nelec)
    t = torch.from_numpy(vec).cuda()
    out = ffsim.apply_orbital_rotation(t, u, norb, nelec)
    assert out.is_cuda and out.dtype == torch.complex128 and out.data_ptr() != t.data_ptr()
    assert rel_err(out.cpu().numpy(), want) < TOL
    np.testing.assert_array_equal(t.cpu().numpy(), vec)  # copy=True leaves the input alone
    out2 = ffsim.apply_orbital_rotation(t, u, norb, nelec, copy=False)
    assert out2.data_ptr() == t.data_ptr()  # copy=False works in place on the device
    assert rel_err(t.cpu().numpy(), want) < TOL
    cpu_t = ffsim.apply_orbital_rotation(torch.from_numpy(vec), u, norb, nelec)
    assert not cpu_t.is_cuda and rel_err(cpu_t.numpy(), want) < TOL
    with pytest.raises(ValueError):
        ffsim.apply_orbital_rotation(vec[:-1], u, norb, nelec)


# ------------------------------------------------------------------ diagonal operators

@pytest.mark.parametrize("norb,nelec", NORB_NELEC_CASES + MEDIUM_CASES)
@pytest.mark.parametrize("z_rep", [False, True])
def test_diag_coulomb_evolution(norb, nelec, z_rep):
    rng = np.random.default_rng(norb + 31)
    vec = _state(norb, nelec, rng)
    maa = rand.random_real_symmetric_matrix(norb, seed=rng)
    mab = rng.standard_normal((norb, norb))  # non-symmetric alpha-beta (diag_coulomb_test.py:159)
    mbb = rand.random_real_symmetric_matrix(norb, seed=rng)
    for mat in (maa, (maa, mab, mbb), (None, mab, None), (maa, None, mbb), (None, None, None), np.stack([maa, mab, mbb])):
        got = ffsim.apply_diag_coulomb_evolution(vec, mat, 0.7, norb, nelec, z_representation=z_rep)
        omat = mat if not isinstance(mat, np.ndarray) or mat.ndim == 2 else tuple(mat)
        want = gates.apply_diag_coulomb_evolution(vec, omat, 0.7, norb, nelec, z_representation=z_rep)
        assert rel_err(got, want) < TOL


@pytest.mark.parametrize("norb,nelec", [(4, (2, 2)), (6, (3, 2)), (9, (4, 4))])
def test_diag_coulomb_evolution_rotated_and_spinless(norb, nelec):
    rng = np.random.default_rng(12)
    vec = _state(norb, nelec, rng)
    mat = rand.random_real_symmetric_matrix(norb, seed=rng)
    ua, ub = rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)
    for rot in (ua, (ua, ub), (None, ub)):
        got = ffsim.apply_diag_coulomb_evolution(vec, mat, 0.4, norb, nelec, orbital_rotation=rot)
        want = gates.apply_diag_coulomb_evolution(vec, mat, 0.4, norb, nelec, orbital_rotation=rot)
        assert rel_err(got, want) < TOL
    v1 = _state(norb, nelec[0], rng)
    got = ffsim.apply_diag_coulomb_evolution(v1, mat, 0.4, norb, nelec[0], orbital_rotation=ua)
    want = gates.apply_diag_coulomb_evolution(v1, mat, 0.4, norb, nelec[0], orbital_rotation=(ua, None))
    assert rel_err(got, want) < TOL
    with pytest.raises(NotImplementedError):
        ffsim.apply_diag_coulomb_evolution(v1, mat, 0.4, norb, nelec[0], z_representation=True)


@pytest.mark.parametrize("norb,nelec", NORB_NELEC_CASES + MEDIUM_CASES)
def test_num_op_sum_evolution(norb, nelec):
    rng = np.random.default_rng(norb + 77)
    vec = _state(norb, nelec, rng)
    ca, cb = rng.standard_normal(norb), rng.standard_normal(norb)
    u = rand.random_unitary(norb, seed=rng)
    for coeffs in (ca, (ca, cb), (ca, None), (None, cb)):
        got = ffsim.apply_num_op_sum_evolution(vec, coeffs, 0.9, norb, nelec)
        assert rel_err(got, gates.apply_num_op_sum_evolution(vec, coeffs, 0.9, norb, nelec)) < TOL
    got = ffsim.apply_num_op_sum_evolution(vec, ca, 0.9, norb, nelec, orbital_rotation=u)
    assert rel_err(got, gates.apply_num_op_sum_evolution(vec, ca, 0.9, norb, nelec, orbital_rotation=u)) < TOL
    v1 = _state(norb, nelec[0], rng)
    got = ffsim.apply_num_op_sum_evolution(v1, ca, 0.9, norb, nelec[0], orbital_rotation=u)
    assert rel_err(got, gates.apply_num_op_sum_evolution(v1, ca, 0.9, norb, nelec[0], orbital_rotation=u)) < TOL


@pytest.mark.parametrize("norb,nelec", NORB_NELEC_CASES + MEDIUM_CASES)
@pytest.mark.parametrize("z_rep", [False, True])
def test_contract_diag_coulomb(norb, nelec, z_rep):
    rng = np.random.default_rng(norb + 5)
    vec = _state(norb, nelec, rng)
    maa = rand.random_real_symmetric_matrix(norb, seed=rng)
    mab = rng.standard_normal((norb, norb))
    mbb = rand.random_real_symmetric_matrix(norb, seed=rng)
    for mat in (maa, (maa, mab, mbb), (None, mab, None), (maa, None, None)):
        got = ffsim.contract_diag_coulomb(vec, mat, norb, nelec, z_representation=z_rep)
        want = contract.contract_diag_coulomb(vec, mat, norb, nelec, z_representation=z_rep)
        assert np.linalg.norm(got - want) <= TOL * max(np.linalg.norm(want), 1.0)
    u = rand.random_unitary(norb, seed=rng)
    linop = ffsim.contract.diag_coulomb_linop(maa, norb, nelec, orbital_rotation=u, z_representation=z_rep)
    want = contract.diag_coulomb_matvec(vec, maa, norb, nelec, orbital_rotation=u, z_representation=z_rep)
    assert np.linalg.norm(linop @ vec - want) <= TOL * max(np.linalg.norm(want), 1.0)


@pytest.mark.parametrize("norb,nelec", NORB_NELEC_CASES + MEDIUM_CASES)
def test_contract_num_op_sum(norb, nelec):
    rng = np.random.default_rng(norb + 6)
    vec = _state(norb, nelec, rng)
    coeffs = rng.standard_normal(norb)
    got = ffsim.contract_num_op_sum(vec, coeffs, norb, nelec)
    want = contract.contract_num_op_sum(vec, coeffs, norb, nelec)
    assert np.linalg.norm(got - want) <= TOL * max(np.linalg.norm(want), 1.0)
    u = rand.random_unitary(norb, seed=rng)
    linop = ffsim.contract.num_op_sum_linop(coeffs, norb, nelec, orbital_rotation=u)
    want = contract.num_op_sum_matvec(vec, coeffs, norb, nelec, orbital_rotation=u)
    assert np.linalg.norm(linop @ vec - want) <= TOL * max(np.linalg.norm(want), 1.0)


# ------------------------------------------------------------------ drivers

@pytest.mark.parametrize("norb,nelec,n_reps,final", [(4, (2, 2), 1, False), (6, (3, 2), 2, True), (8, (4, 4), 3, False)])
def test_ucj_spin_balanced(norb, nelec, n_reps, final):
    op = ffsim.random.random_ucj_op_spin_balanced(norb, n_reps=n_reps, with_final_orbital_rotation=final, seed=norb)
    rng = np.random.default_rng(9)
    vec = _state(norb, nelec, rng)
    got = ffsim.apply_unitary(vec, op, norb=norb, nelec=nelec)
    want = models.ucj_spin_balanced_apply(vec, op.diag_coulomb_mats, op.orbital_rotations,
                                          op.final_orbital_rotation, norb, nelec)
    assert rel_err(got, want) < TOL
    with pytest.raises(TypeError):
        ffsim.apply_unitary(vec[: math.comb(norb, nelec[0])], op, norb=norb, nelec=nelec[0])
    with pytest.raises(TypeError):
        ffsim.apply_unitary(vec, object(), norb=norb, nelec=nelec)


def test_ucj_validation_errors():
    op = ffsim.random.random_ucj_op_spin_balanced(4, n_reps=2, seed=1)
    with pytest.raises(ValueError, match="shape"):
        ffsim.UCJOpSpinBalanced(op.diag_coulomb_mats[:, 0], op.orbital_rotations)
    with pytest.raises(ValueError, match="unitary"):
        ffsim.UCJOpSpinBalanced(op.diag_coulomb_mats, op.orbital_rotations * 1.1)
    bad = op.diag_coulomb_mats.copy()
    bad[0, 0, 0, 1] += 1
    with pytest.raises(ValueError, match="symmetric"):
        ffsim.UCJOpSpinBalanced(bad, op.orbital_rotations)
    ffsim.UCJOpSpinBalanced(bad, op.orbital_rotations, validate=False)


def test_lucj_c1_shape_against_c_oracle():
    """BASELINE config C1: LUCJ n_reps=2 on Hartree-Fock, norb=12, nelec=(6,6)."""
    norb, nelec = 12, (6, 6)
    pairs_aa = [(p, p + 1) for p in range(norb - 1)]
    pairs_ab = [(p, p) for p in range(norb)]
    op = ffsim.random.random_ucj_op_spin_balanced(norb, n_reps=2, interaction_pairs=(pairs_aa, pairs_ab), seed=1201)
    vec = ffsim.hartree_fock_state(norb, nelec)
    got = ffsim.apply_unitary(vec, op, norb=norb, nelec=nelec)
    want = cref.ucj_spin_balanced_apply(vec, op.diag_coulomb_mats, op.orbital_rotations, None, norb, nelec)
    assert rel_err(got, want) < TOL
    assert abs(np.linalg.norm(got) - 1) < 1e-12


@pytest.mark.parametrize("order,n_steps,z_rep", [(0, 1, False), (1, 2, False), (2, 1, True), (0, 3, True)])
def test_trotter_double_factorized(order, n_steps, z_rep):
    norb, nelec = 5, (3, 2)
    ham = ffsim.random.random_double_factorized_hamiltonian(norb, rank=4, z_representation=z_rep, seed=21)
    rng = np.random.default_rng(22)
    vec = _state(norb, nelec, rng)
    got = ffsim.simulate_trotter_double_factorized(vec, ham, 0.3, norb=norb, nelec=nelec, n_steps=n_steps, order=order)
    want = models.simulate_trotter_double_factorized(
        vec, ham.one_body_tensor, ham.diag_coulomb_mats, ham.orbital_rotations, ham.constant, z_rep, 0.3,
        norb=norb, nelec=nelec, n_steps=n_steps, order=order)
    assert rel_err(got, want) < TOL
    with pytest.raises(ValueError):
        ffsim.simulate_trotter_double_factorized(vec, ham, 0.3, norb=norb, nelec=nelec, order=-1)
    with pytest.raises(ValueError):
        ffsim.simulate_trotter_double_factorized(vec, ham, 0.3, norb=norb, nelec=nelec, n_steps=-1)
    same = ffsim.simulate_trotter_double_factorized(vec, ham, 0.3, norb=norb, nelec=nelec, n_steps=0)
    np.testing.assert_array_equal(same, vec)


def test_diagonal_coulomb_hamiltonian_linear_operator():
    norb, nelec = 6, (3, 2)
    ham = ffsim.random.random_diagonal_coulomb_hamiltonian(norb, seed=33)
    rng = np.random.default_rng(34)
    vec = _state(norb, nelec, rng)
    linop = ffsim.linear_operator(ham, norb=norb, nelec=nelec)
    want = models.diagonal_coulomb_hamiltonian_matvec(vec, ham.one_body_tensor, ham.diag_coulomb_mats, ham.constant, norb, nelec)
    got = linop @ vec
    assert np.linalg.norm(got - want) <= TOL * np.linalg.norm(want)
    assert np.linalg.norm(linop.matvec(vec) - want) <= TOL * np.linalg.norm(want)
    import torch

    t = torch.from_numpy(vec).cuda()
    out = linop @ t
    assert out.is_cuda and np.linalg.norm(out.cpu().numpy() - want) <= TOL * np.linalg.norm(want)
    np.testing.assert_array_equal(t.cpu().numpy(), vec)
    energy = np.vdot(vec, got).real
    assert abs(energy - np.vdot(vec, want).real) < 1e-11
    with pytest.raises(TypeError):
        ffsim.linear_operator(object(), norb=norb, nelec=nelec)


def test_docs_hubbard_ground_energy_and_split_op_fidelities():
    # docs/explanations/diag-coulomb-hamiltonian.ipynb cells 5, 7
    norb, nelec = 4, (2, 2)
    h = np.array([[-2, -1, -1, 0], [-1, -2, 0, -1], [-1, 0, -2, -1], [0, -1, -1, -2]], dtype=complex)
    ham = ffsim.DiagonalCoulombHamiltonian(h, np.stack([np.zeros((4, 4)), 4.0 * np.eye(4)]), constant=0)
    linop = ffsim.linear_operator(ham, norb=norb, nelec=nelec)
    eigs, _ = scipy.sparse.linalg.eigsh(linop, k=1, which="SA")
    assert eigs[0] == pytest.approx(-10.10274848346205, abs=1e-10)
    dim = ffsim.dim(norb, nelec)
    dense = np.stack([linop @ e for e in np.eye(dim, dtype=complex)], axis=1)
    vec = ffsim.hartree_fock_state(norb, nelec)
    exact = scipy.linalg.expm(-1j * dense) @ vec
    for n_steps, fid in {1: 0.45702529, 2: 0.95880093, 5: 0.99915103, 10: 0.99994861}.items():
        res = ffsim.simulate_trotter_diag_coulomb_split_op(vec, ham, 1.0, norb=norb, nelec=nelec, n_steps=n_steps, order=1)
        assert abs(np.vdot(res, exact)) == pytest.approx(fid, abs=5e-9)


def test_double_factorized_hamiltonian_linear_operator():
    norb, nelec = 5, (2, 3)
    for z_rep in (False, True):
        ham = ffsim.random.random_double_factorized_hamiltonian(norb, rank=3, z_representation=z_rep, seed=41)
        rng = np.random.default_rng(42)
        vec = _state(norb, nelec, rng)
        got = ffsim.linear_operator(ham, norb=norb, nelec=nelec) @ vec
        want = models.double_factorized_hamiltonian_matvec(
            vec, ham.one_body_tensor, ham.diag_coulomb_mats, ham.orbital_rotations, ham.constant, z_rep, norb, nelec)
        assert np.linalg.norm(got - want) <= TOL * np.linalg.norm(want)


# ------------------------------------------------------------------ SURVEY.md 8f rows
@pytest.mark.parametrize("norb, nelec", [(6, (3, 2)), (7, (2, 4))])
def test_ucj_spin_unbalanced_and_spinless(norb, nelec):
    rng = np.random.default_rng(51)
    n_reps = 2
    mats3 = np.stack([np.stack([rand.random_real_symmetric_matrix(norb, seed=rng), rng.standard_normal((norb, norb)),
                                rand.random_real_symmetric_matrix(norb, seed=rng)]) for _ in range(n_reps)])
    rots2 = np.stack([np.stack([rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)])
                      for _ in range(n_reps)])
    final2 = np.stack([rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)])
    vec = _state(norb, nelec, rng)
    for final in (None, final2):
        op = ffsim.UCJOpSpinUnbalanced(mats3, rots2, final_orbital_rotation=final)
        got = ffsim.apply_unitary(vec, op, norb=norb, nelec=nelec)
        want = models.ucj_spin_unbalanced_apply(vec, mats3, rots2, final, norb, nelec)
        assert rel_err(got, want) <= TOL
    # spinless operator on a spinful state and on a spinless one
    op = ffsim.UCJOpSpinless(mats3[:, 0], rots2[:, 0], final_orbital_rotation=final2[0])
    got = ffsim.apply_unitary(vec, op, norb=norb, nelec=nelec)
    assert rel_err(got, models.ucj_spinless_apply(vec, mats3[:, 0], rots2[:, 0], final2[0], norb, nelec)) <= TOL
    nocc = nelec[0]
    vec1 = rand.random_state_vector(math.comb(norb, nocc), seed=rng)
    got = ffsim.apply_unitary(vec1, op, norb=norb, nelec=nocc)
    assert rel_err(got, models.ucj_spinless_apply(vec1, mats3[:, 0], rots2[:, 0], final2[0], norb, nocc)) <= TOL


def test_qdrift_double_factorized():
    import torch

    norb, nelec = 5, (2, 3)
    ham = ffsim.random.random_double_factorized_hamiltonian(norb, rank=3, seed=61)
    vec = _state(norb, nelec, np.random.default_rng(62))
    args = (ham.one_body_tensor, ham.diag_coulomb_mats, ham.orbital_rotations, False)
    probs = np.array([0.1, 0.2, 0.3, 0.4])
    for symmetric in (False, True):
        for p in ("norm", "uniform", probs):
            got = ffsim.simulate_qdrift_double_factorized(vec, ham, 0.4, norb=norb, nelec=nelec, n_steps=6,
                                                          symmetric=symmetric, probabilities=p, n_samples=3, seed=7)
            want = models.simulate_qdrift_double_factorized(vec, *args, 0.4, norb=norb, nelec=nelec, n_steps=6,
                                                            symmetric=symmetric, probabilities=p, n_samples=3, seed=7)
            assert got.shape == (3, vec.size) and rel_err(got, want) <= TOL
    assert np.array_equal(probs, [0.1, 0.2, 0.3, 0.4]), "the caller's probabilities must not be modified"
    # zero steps / zero time return copies of the input; CUDA tensor in -> CUDA tensor out
    out = ffsim.simulate_qdrift_double_factorized(vec, ham, 0.4, norb=norb, nelec=nelec, n_steps=0)
    assert np.array_equal(out, vec) and out is not vec
    dev = torch.from_numpy(vec).cuda()
    out = ffsim.simulate_qdrift_double_factorized(dev, ham, 0.4, norb=norb, nelec=nelec, n_steps=3, seed=9)
    want = models.simulate_qdrift_double_factorized(vec, *args, 0.4, norb=norb, nelec=nelec, n_steps=3, seed=9)
    assert out.is_cuda and torch.equal(dev.cpu(), torch.from_numpy(vec)) and rel_err(out.cpu().numpy(), want) <= TOL
    with pytest.raises(ValueError, match="n_steps"):
        ffsim.simulate_qdrift_double_factorized(vec, ham, 0.4, norb=norb, nelec=nelec, n_steps=-1)
    with pytest.raises(ValueError, match="n_samples"):
        ffsim.simulate_qdrift_double_factorized(vec, ham, 0.4, norb=norb, nelec=nelec, n_samples=0)
    # state-dependent probabilities (Wick expectations in the Hartree-Fock determinant): same trajectory as
    # passing those probabilities explicitly
    rdm = np.zeros((2 * norb, 2 * norb))
    for i in list(range(nelec[0])) + [norb + j for j in range(nelec[1])]:
        rdm[i, i] = 1.0
    probs_opt = ffsim.qdrift_probabilities(ham, "optimal", nelec=nelec, one_rdm=rdm)
    assert abs(probs_opt.sum() - 1) < 1e-12 and (probs_opt >= 0).all()
    got = ffsim.simulate_qdrift_double_factorized(vec, ham, 0.4, norb=norb, nelec=nelec, n_steps=4,
                                                  probabilities="optimal", one_rdm=rdm, seed=11)
    want = ffsim.simulate_qdrift_double_factorized(vec, ham, 0.4, norb=norb, nelec=nelec, n_steps=4,
                                                   probabilities=probs_opt, seed=11)
    assert rel_err(got, want) <= TOL
    with pytest.raises(ValueError, match="requires one_rdm"):
        ffsim.simulate_qdrift_double_factorized(vec, ham, 0.4, norb=norb, nelec=nelec, probabilities="optimal")


# ------------------------------------------------------------------ host-resident states (streamed copies)

@pytest.mark.parametrize("opts", [{}, {"smem_bytes": 16 * 1024, "beta_mode": 2}])
@pytest.mark.parametrize("n_chunks", [1, 3, 7])
def test_evolve_host_matches_sequential_calls(opts, n_chunks):
    """ffsim_b200.evolve_host (column strips in, row blocks out, kernels overlapped with the copies) against
    the same public functions called one after the other, single-pass and multi-pass/transposed plans."""
    for k, v in opts.items():
        _lib.set_option(k, v)
    norb, nelec = 10, (5, 4)
    rng = np.random.default_rng(1010)
    vec = _state(norb, nelec, rng)
    u1, u2 = rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)
    mat = rand.random_real_symmetric_matrix(norb, seed=rng)
    coeffs = rng.standard_normal(norb)
    pinned = ffsim.pinned_empty(vec.size)
    pinned[:] = vec
    cases = [
        [("orbital_rotation", u1), ("diag_coulomb", mat, 0.3)],                       # the bench's step
        [("diag_coulomb", mat, 0.3, True)],                                           # everything column-local
        [("orbital_rotation", (u1, None))], [("orbital_rotation", (None, u2))],
        [("orbital_rotation", u1), ("diag_coulomb", (mat, mat + 0.2, None), 0.1), ("orbital_rotation", (u2, u1)),
         ("num_op_sum", coeffs, 0.7)],
        [("num_op_sum", (coeffs, None), 0.2), ("orbital_rotation", u2)],
    ]
    for steps in cases:
        want = vec
        for step in steps:
            if step[0] == "orbital_rotation":
                want = ffsim.apply_orbital_rotation(want, step[1], norb, nelec)
            elif step[0] == "diag_coulomb":
                want = ffsim.apply_diag_coulomb_evolution(want, step[1], step[2], norb, nelec,
                                                          z_representation=len(step) > 3 and step[3])
            else:
                want = ffsim.apply_num_op_sum_evolution(want, step[1], step[2], norb, nelec)
        for src in (pinned, vec):
            before = src.copy()
            got = ffsim.evolve_host(src, steps, norb, nelec, n_chunks=n_chunks)
            assert np.array_equal(src, before), "the input must not be modified"
            assert rel_err(got, want) <= TOL, (steps[0][0], len(steps), n_chunks, rel_err(got, want))
    with pytest.raises(ValueError, match="unknown step"):
        ffsim.evolve_host(vec, [("rotate", u1)], norb, nelec)
    with pytest.raises(ValueError, match="entries"):
        ffsim.evolve_host(vec[:-1], [("orbital_rotation", u1)], norb, nelec)


@pytest.mark.parametrize("opts", [{}, {"smem_bytes": 16 * 1024, "beta_mode": 2}])
def test_evolve_host_async_overlapped_applications(opts):
    """Applications started back to back (ffsim_b200.evolve_host_async / evolve_host_many: ring of two device
    buffers, uploads of one application under the kernels and downloads of the one before) give, each, what
    a lone evolve_host call gives -- different states, different sizes in between, inputs untouched."""
    for k, v in opts.items():
        _lib.set_option(k, v)
    norb, nelec = 10, (4, 5)
    rng = np.random.default_rng(2020)
    u = rand.random_unitary(norb, seed=rng)
    mat = rand.random_real_symmetric_matrix(norb, seed=rng)
    steps = [("orbital_rotation", u), ("diag_coulomb", mat, 0.25)]
    vecs = []
    for _ in range(7):
        pinned = ffsim.pinned_empty(ffsim.dim(norb, nelec))
        pinned[:] = _state(norb, nelec, rng)
        vecs.append(pinned)
    want = [ffsim.apply_diag_coulomb_evolution(ffsim.apply_orbital_rotation(v, u, norb, nelec), mat, 0.25, norb, nelec)
            for v in vecs]
    before = [v.copy() for v in vecs]
    got = ffsim.evolve_host_many(vecs, steps, norb, nelec, n_chunks=3)
    assert len(got) == len(vecs)
    for g, w, v, b in zip(got, want, vecs, before):
        assert rel_err(g, w) <= TOL
        assert np.array_equal(v, b)
    # handles: results may be collected in any order; a different state size in between re-sizes the ring
    handles = [ffsim.evolve_host_async(v, steps, norb, nelec, n_chunks=2) for v in vecs[:4]]
    small = _state(6, (3, 3), rng)
    u6 = rand.random_unitary(6, seed=rng)
    h_small = ffsim.evolve_host_async(small, [("orbital_rotation", u6)], 6, (3, 3))
    assert rel_err(h_small.result(), ffsim.apply_orbital_rotation(small, u6, 6, (3, 3))) <= TOL
    for k in (3, 0, 2, 1):
        assert rel_err(handles[k].result(), want[k]) <= TOL
        assert handles[k].done()
    ffsim.release_device_buffers()


# ------------------------------------------------------------------ BASELINE shapes: properties

def test_c2_shape_against_c_oracle_and_properties():
    """BASELINE config C2: norb=16, nelec=(5,5), 19.1M amplitudes."""
    import torch

    norb, nelec = 16, (5, 5)
    rng = np.random.default_rng(1602)
    dim = ffsim.dim(norb, nelec)
    vec = rand.random_state_vector(dim, seed=rng)
    u = rand.random_unitary(norb, seed=rng)
    mat = rand.random_real_symmetric_matrix(norb, seed=rng)
    got = ffsim.apply_orbital_rotation(vec, u, norb, nelec)
    want = cref.apply_orbital_rotation(vec, u, norb, nelec)
    assert rel_err(got, want) < TOL
    got = ffsim.apply_diag_coulomb_evolution(vec, mat, 1.0, norb, nelec)
    assert rel_err(got, cref.apply_diag_coulomb_evolution(vec, mat, 1.0, norb, nelec)) < TOL
    # size-independent properties on the device: unitarity and inverse
    t = torch.from_numpy(vec).cuda()
    r = ffsim.apply_orbital_rotation(t, u, norb, nelec)
    assert abs(torch.linalg.vector_norm(r).item() - 1) < 1e-12
    back = ffsim.apply_orbital_rotation(r, u.T.conj(), norb, nelec)
    assert (torch.linalg.vector_norm(back - t) / torch.linalg.vector_norm(t)).item() < TOL
    # Hartree-Fock image = outer product of Slater minors (closed form, not Givens based)
    hf = ffsim.hartree_fock_state(norb, nelec)
    minors = compound.slater_minors(u, norb, 5)
    assert rel_err(ffsim.apply_orbital_rotation(hf, u, norb, nelec), np.outer(minors, minors).reshape(-1)) < TOL


def test_plan_is_fused():
    """The whole point: a few passes over the state instead of n(n-1)/2 per spin."""
    norb, nelec = 16, (5, 5)
    u = rand.random_unitary(norb, seed=3)
    plan = get_plan(norb, nelec, u, u)
    assert plan.n_state_passes() <= 10, plan.describe()
