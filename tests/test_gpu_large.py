"""GPU parity at the BASELINE shapes that are too large for a whole-state CPU run.

An orbital rotation of one spin acts on every column (row) of the state independently, so the CPU
oracle (oracle/cref.py) can check a full-size GPU run exactly on a sample of columns (rows): the
sampled slab of the INPUT goes through the oracle and is compared with the same slab of the GPU
OUTPUT.  Every sampled column passes through every sweep, tile and register block of the multi-pass
plan, and the samples are spread over first / last / interior tiles.  Diagonal operators are checked
entry by entry on sampled (row, column) pairs from the closed form.  Tolerance: relative 2-norm
<= 1e-12 (BASELINE.json north_star).

Covers VERDICT round 1 "pin parity at the big shapes": C3 shape (norb=18, nelec=(7,7): 3 sweeps per
side, beta side on a transposed copy), the forced multi-pass + transposed plan at norb=16 (8,8), the
norb=14 (6,6) twin of C4 (LUCJ n_reps=3, seed 2004) end to end, and a DF-Trotter step at the twin shape.
"""

import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ffsim_b200 as ffsim  # noqa: E402
from ffsim_b200 import _lib  # noqa: E402
from ffsim_b200.gates.orbital_rotation import get_plan  # noqa: E402
from oracle import cistring, cref, gates, givens, models, rand  # noqa: E402

TOL = 1e-12


def rel_err(got, want):
    n = np.linalg.norm(want)
    return np.linalg.norm(np.asarray(got) - want) / (n if n > 0 else 1.0)


@pytest.fixture(autouse=True)
def _default_options():
    saved = {k: _lib.get_option(k) for k in ("smem_bytes", "min_cols", "max_cols", "sub_window", "threads", "beta_mode", "bulk_copies")}
    yield
    for k, v in saved.items():
        _lib.set_option(k, v)


def _random_device_state(dim, seed):
    import torch

    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    vec = torch.empty(dim, dtype=torch.complex128, device="cuda")
    torch.view_as_real(vec).normal_(generator=gen)
    vec.mul_(1.0 / float(torch.linalg.vector_norm(vec)))
    return vec


def _sample(n, k, rng):
    """k indices of range(n): both ends, the neighbours of a few tile boundaries, the rest random."""
    fixed = [0, 1, 2, 7, 8, n // 2 - 1, n // 2, n - 9, n - 8, n - 2, n - 1]
    idx = set(i for i in fixed if 0 <= i < n)
    while len(idx) < min(k, n):
        idx.add(int(rng.integers(n)))
    return np.array(sorted(idx))


def _check_sides_on_slabs(norb, nelec, seed, n_samples=24):
    """alpha-only and beta-only rotations of a full-size random state, each checked on sampled slabs;
    then both spins at once, checked through the commuting structure (alpha-only o beta-only)."""
    import torch

    dim_a, dim_b = math.comb(norb, nelec[0]), math.comb(norb, nelec[1])
    rng = np.random.default_rng(seed)
    ua, ub = rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)
    vec = _random_device_state(dim_a * dim_b, seed)
    cols = _sample(dim_b, n_samples, rng)
    rows = _sample(dim_a, n_samples, rng)
    cols_t, rows_t = torch.from_numpy(cols).cuda(), torch.from_numpy(rows).cuda()
    in_cols = vec.view(dim_a, dim_b)[:, cols_t].cpu().numpy()  # [dim_a x k]
    in_rows = vec.view(dim_a, dim_b)[rows_t, :].cpu().numpy()  # [k x dim_b]

    # alpha side: every column independently
    out = ffsim.apply_orbital_rotation(vec, (ua, None), norb, nelec, copy=True)
    want = np.ascontiguousarray(in_cols)
    cref._rotate_one_spin(want, givens.givens_decomposition(ua), norb, nelec[0])
    got = out.view(dim_a, dim_b)[:, cols_t].cpu().numpy()
    err_a = rel_err(got, want)
    assert err_a < TOL, ("alpha", norb, nelec, err_a)
    assert abs(float(torch.linalg.vector_norm(out)) - 1.0) < 1e-12
    del out

    # beta side: every row independently (the oracle rotates the transposed slab, like the reference)
    out = ffsim.apply_orbital_rotation(vec, (None, ub), norb, nelec, copy=True)
    want = np.ascontiguousarray(in_rows.T)
    cref._rotate_one_spin(want, givens.givens_decomposition(ub), norb, nelec[1])
    got = out.view(dim_a, dim_b)[rows_t, :].cpu().numpy()
    err_b = rel_err(got, want.T)
    assert err_b < TOL, ("beta", norb, nelec, err_b)
    assert abs(float(torch.linalg.vector_norm(out)) - 1.0) < 1e-12

    # both spins in one call == alpha-only applied to the beta-only result (the sides commute); the
    # alpha-only and beta-only paths were pinned above, so this pins the combined plan at full size
    both = ffsim.apply_orbital_rotation(vec, (ua, ub), norb, nelec, copy=True)
    del vec
    ffsim.apply_orbital_rotation(out, (ua, None), norb, nelec, copy=False)
    diff = float(torch.linalg.vector_norm(both - out))
    assert diff < TOL, ("both", norb, nelec, diff)
    return err_a, err_b, diff


def test_orbital_rotation_c3_shape_sampled_slabs():
    """BASELINE configs[2] shape: norb=18, nelec=(7,7), 16.2 GB; three sweeps per side, beta transposed."""
    norb, nelec = 18, (7, 7)
    u = rand.random_unitary(norb, seed=3)
    plan = get_plan(norb, nelec, u, u)
    desc = plan.describe()
    assert plan.n_state_passes() >= 6, desc  # a genuine multi-pass plan
    _check_sides_on_slabs(norb, nelec, seed=1803)


def test_orbital_rotation_forced_multipass_transposed_norb16():
    """norb=16, nelec=(8,8) (2.65 GB) with a small shared-memory budget: several sweeps per side and
    the beta side on a transposed copy -- the plan shape C3/C4 use, forced at a mid-size state."""
    norb, nelec = 16, (8, 8)
    _lib.set_option("smem_bytes", 96 * 1024)
    _lib.set_option("beta_mode", 2)
    u = rand.random_unitary(norb, seed=5)
    desc = get_plan(norb, nelec, u, u).describe()
    assert "layout=transposed" in desc and desc.count("[lo=") >= 6, desc
    _check_sides_on_slabs(norb, nelec, seed=1608)


def test_diag_coulomb_c3_shape_sampled_entries():
    """Diagonal Coulomb evolution (number and Z representation) at the C3 shape against the closed form
    exp(-i t sum ...) on sampled amplitudes (src/gates/diag_coulomb.rs:21,95 restated in oracle/gates.py)."""
    import torch

    norb, nelec = 18, (7, 7)
    dim_a, dim_b = math.comb(norb, nelec[0]), math.comb(norb, nelec[1])
    rng = np.random.default_rng(77)
    mat_aa = rand.random_real_symmetric_matrix(norb, seed=rng)
    mat_ab = rng.standard_normal((norb, norb))  # alpha-beta block need not be symmetric
    mat_bb = rand.random_real_symmetric_matrix(norb, seed=rng)
    time = 0.37
    vec = _random_device_state(dim_a * dim_b, 99)
    rows, cols = _sample(dim_a, 40, rng), _sample(dim_b, 40, rng)
    rows_t, cols_t = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda()
    before = vec.view(dim_a, dim_b)[rows_t][:, cols_t].cpu().numpy()
    strings_a = cistring.make_strings(range(norb), nelec[0])
    strings_b = cistring.make_strings(range(norb), nelec[1])
    for z_rep in (False, True):
        out = ffsim.apply_diag_coulomb_evolution(vec, (mat_aa, mat_ab, mat_bb), time, norb, nelec,
                                                 z_representation=z_rep, copy=True)
        got = out.view(dim_a, dim_b)[rows_t][:, cols_t].cpu().numpy()
        del out
        want = np.empty_like(before)
        for i, a in enumerate(rows):
            na = np.array([(int(strings_a[a]) >> p) & 1 for p in range(norb)], dtype=float)
            for j, b in enumerate(cols):
                nb = np.array([(int(strings_b[b]) >> p) & 1 for p in range(norb)], dtype=float)
                if z_rep:
                    za, zb = 1 - 2 * na, 1 - 2 * nb
                    e = 0.25 * (0.5 * (za @ mat_aa @ za - np.trace(mat_aa)) + 0.5 * (zb @ mat_bb @ zb - np.trace(mat_bb))
                                + za @ mat_ab @ zb)
                    # the reference's z representation keeps the j == k terms of the same-spin blocks out
                    # (src/gates/diag_coulomb.rs:117-135: pairs j < k only) and counts alpha-beta over all pairs
                else:
                    e = 0.5 * (na @ mat_aa @ na) + 0.5 * (nb @ mat_bb @ nb) + na @ mat_ab @ nb
                want[i, j] = before[i, j] * np.exp(-1j * time * e)
        err = rel_err(got, want)
        assert err < TOL, ("diag", z_rep, err)


def test_lucj_c4_twin_norb14():
    """SURVEY.md section 8d, C4 check 2: the scaled-down twin of the 254 GB LUCJ run -- norb=14,
    nelec=(6,6), n_reps=3, the same seed (2004) -- end to end against the C oracle."""
    norb, nelec = 14, (6, 6)
    rng = np.random.default_rng(2004)
    op = ffsim.random.random_ucj_op_spin_balanced(norb, n_reps=3, with_final_orbital_rotation=True, seed=rng)
    vec = models.hartree_fock_state(norb, nelec)
    got = ffsim.apply_unitary(vec, op, norb=norb, nelec=nelec)
    want = cref.ucj_spin_balanced_apply(vec, op.diag_coulomb_mats, op.orbital_rotations, op.final_orbital_rotation,
                                        norb, nelec)
    assert rel_err(got, want) < TOL
    # and with the multi-pass + transposed plan the big shapes use
    _lib.set_option("smem_bytes", 64 * 1024)
    _lib.set_option("beta_mode", 2)
    got2 = ffsim.apply_unitary(vec, op, norb=norb, nelec=nelec)
    assert rel_err(got2, want) < TOL


def test_trotter_step_twin_norb14_multipass():
    """One double-factorized Trotter step (order 1, rank 3) at norb=14 (6,6) with the plan shape of C3
    (several sweeps per side, transposed beta side) against the numpy/C oracle."""
    norb, nelec = 14, (6, 6)
    _lib.set_option("smem_bytes", 64 * 1024)
    _lib.set_option("beta_mode", 2)
    df = ffsim.random.random_double_factorized_hamiltonian(norb, rank=3, seed=1403)
    rng = np.random.default_rng(5)
    vec = rand.random_state_vector(models.dim(norb, nelec), seed=rng)
    got = ffsim.simulate_trotter_double_factorized(vec, df, 0.25, norb=norb, nelec=nelec, n_steps=1, order=1)
    # the oracle with the C kernels for the rotations: same driver, python/ffsim/trotter/double_factorized.py:25-125
    import scipy.linalg

    want = vec.copy()
    basis = np.eye(norb, dtype=complex)
    for term, t in models.simulate_trotter_step_iterator(1 + len(df.diag_coulomb_mats), 0.25, 1):
        if term == 0:
            basis = scipy.linalg.expm(-1j * t * df.one_body_tensor) @ basis
        else:
            u = df.orbital_rotations[term - 1]
            want = cref.apply_orbital_rotation(want, u.T.conj() @ basis, norb, nelec, copy=False)
            want = cref.apply_diag_coulomb_evolution(want, df.diag_coulomb_mats[term - 1], t, norb, nelec, copy=False)
            basis = u
    want = cref.apply_orbital_rotation(want, basis, norb, nelec, copy=False)
    want = want * np.exp(-1j * 0.25 * df.constant)
    assert rel_err(got, want) < TOL


def _wick_energy(one_body, jaa, jab, constant, u, n_alpha, n_beta):
    """<H> of the determinant U|HF> from its one-body density matrices (closed form, O(norb^2)):
    SURVEY.md section 8d, the C5 check."""
    pa = u[:, :n_alpha] @ u[:, :n_alpha].conj().T
    pb = u[:, :n_beta] @ u[:, :n_beta].conj().T
    e = constant + np.trace(one_body @ pa).real + np.trace(one_body @ pb).real
    for p_ in (pa, pb):  # same spin: <n_p n_q> = P_pp P_qq - |P_pq|^2 (p != q), P_pp (p == q)
        d = np.real(np.diag(p_))
        nn = np.outer(d, d) - np.abs(p_) ** 2
        np.fill_diagonal(nn, d)
        e += 0.5 * np.sum(jaa * nn)
    da, db = np.real(np.diag(pa)), np.real(np.diag(pb))
    e += 0.5 * np.sum(jab * (np.outer(da, db) + np.outer(db, da)))
    return float(e)


@pytest.mark.parametrize("norb,nelec", [(10, (3, 4)), (16, (4, 4)), (20, (3, 3))])
def test_dc_hamiltonian_energy_c5_twins(norb, nelec):
    """BASELINE configs[4] (DiagonalCoulombHamiltonian energy of a rotated determinant, norb=24 (6,6)) at
    shapes one GPU holds: <psi|H|psi> through linear_operator on the device against the closed form from
    the determinant's density matrices -- independent of every kernel and of the oracle."""
    import torch

    ham = ffsim.random.random_diagonal_coulomb_hamiltonian(norb, seed=2405)
    u = ffsim.random.random_unitary(norb, seed=2406)
    state = ffsim.hartree_fock_state(norb, nelec, device="cuda")
    state = ffsim.apply_orbital_rotation(state, u, norb, nelec, copy=False)
    hv = ffsim.linear_operator(ham, norb=norb, nelec=nelec) @ state
    energy = float(torch.vdot(state, hv).real)
    want = _wick_energy(np.asarray(ham.one_body_tensor), *np.asarray(ham.diag_coulomb_mats), ham.constant, u, *nelec)
    assert abs(energy - want) <= 1e-10 * max(1.0, abs(want)), (energy, want)
    assert abs(float(torch.linalg.vector_norm(state)) - 1.0) < 1e-12
