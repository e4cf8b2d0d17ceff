"""Parity against vectors produced by the REFERENCE's own Python code.

tests/golden/reference_vectors.npz was written by tests/golden/make_golden.py, which runs
the reference's drivers (python/ffsim/gates/*.py, contract/*.py, variational/
ucj_spin_balanced.py, trotter/*.py, hamiltonians/diagonal_coulomb_hamiltonian.py) and
generators (random/random.py) over the reference's pure-Python kernel twins
(python/ffsim/_slow/**) -- see tests/golden/ref_shim.py for what had to be substituted.

* not-gpu tests: the CPU oracle (oracle/) reproduces every vector  -> the oracle is pinned;
* gpu tests: the CUDA path (ffsim_b200, through the C ABI) reproduces every vector.

Tolerance: relative 2-norm error <= 1e-12 (BASELINE.json north_star); index tables and the
generator outputs bit-exact.
"""

import collections
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz")
TOL = 1e-12


def _load():
    cases = collections.OrderedDict()
    with np.load(GOLDEN) as z:
        for key in z.files:
            name, field = key.split("::")
            cases.setdefault(name, {})[field] = z[key]
    return cases


CASES = _load()


def names(kind_prefix):
    return [n for n in CASES if n.startswith(kind_prefix)]


def opt(a):
    return None if a.size == 0 and a.ndim == 1 else a


def nelec_of(c):
    ne = c["nelec"]
    return int(ne) if ne.ndim == 0 else (int(ne[0]), int(ne[1]))


def rel_err(got, want):
    got = np.asarray(got)
    assert got.shape == want.shape and got.dtype == np.complex128
    return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300)


class OracleAPI:
    """The CPU oracle behind the reference's signatures."""

    def __init__(self):
        from oracle import contract, gates, models

        self.apply_orbital_rotation = gates.apply_orbital_rotation
        self.apply_diag_coulomb_evolution = gates.apply_diag_coulomb_evolution
        self.apply_num_op_sum_evolution = gates.apply_num_op_sum_evolution
        self.contract_diag_coulomb = contract.contract_diag_coulomb
        self.contract_num_op_sum = contract.contract_num_op_sum
        self.m = models

    def ucj(self, vec, c, norb, nelec):
        return self.m.ucj_spin_balanced_apply(vec, c["diag_coulomb_mats"], c["orbital_rotations"],
                                              opt(c["final_orbital_rotation"]), norb, nelec)

    def trotter_df(self, vec, c, norb, nelec):
        return self.m.simulate_trotter_double_factorized(
            vec, c["one_body_tensor"], c["diag_coulomb_mats"], c["orbital_rotations"], complex(c["constant"]).real,
            bool(c["z"]), float(c["time"]), norb=norb, nelec=nelec, n_steps=int(c["n_steps"]), order=int(c["order"]))

    def dc_matvec(self, vec, c, norb, nelec):
        return self.m.diagonal_coulomb_hamiltonian_matvec(vec, c["one_body_tensor"], c["diag_coulomb_mats"],
                                                          float(c["constant"]), norb, nelec)

    def dc_split_op(self, vec, c, norb, nelec):
        return self.m.simulate_trotter_diag_coulomb_split_op(
            vec, c["one_body_tensor"], c["diag_coulomb_mats"], float(c["constant"]), float(c["time"]), norb=norb,
            nelec=nelec, n_steps=int(c["n_steps"]), order=int(c["order"]))

    def ucj_unbalanced(self, vec, c, norb, nelec):
        return self.m.ucj_spin_unbalanced_apply(vec, c["diag_coulomb_mats"], c["orbital_rotations"],
                                                opt(c["final_orbital_rotation"]), norb, nelec)

    def ucj_spinless(self, vec, c, norb, nelec):
        return self.m.ucj_spinless_apply(vec, c["diag_coulomb_mats"], c["orbital_rotations"],
                                         opt(c["final_orbital_rotation"]), norb, nelec)

    def df_matvec(self, vec, c, norb, nelec):
        return self.m.double_factorized_hamiltonian_matvec(
            vec, c["one_body_tensor"], c["diag_coulomb_mats"], c["orbital_rotations"], float(c["constant"]),
            bool(c["z"]), norb, nelec)

    def qdrift(self, vec, c, norb, nelec):
        return self.m.simulate_qdrift_double_factorized(
            vec, c["one_body_tensor"], c["diag_coulomb_mats"], c["orbital_rotations"], bool(c["z"]), float(c["time"]),
            norb=norb, nelec=nelec, n_steps=int(c["n_steps"]), symmetric=bool(c["symmetric"]),
            probabilities=bytes(c["probabilities"].astype(np.uint8)).decode(), n_samples=int(c["n_samples"]),
            seed=int(c["seed"]))


    def basic(self, vec, c, norb, nelec):
        from oracle import basic_gates as bg

        gate, spin = _text(c["gate"]), _text(c["spin"])
        theta, phi, orbs = float(c["theta"]), float(c["phi"]), tuple(int(x) for x in c["orbs"])
        if gate == "givens":
            return bg.apply_givens_rotation(vec, theta, orbs, norb, nelec, spin, phi)
        if gate == "tunneling":
            return bg.apply_tunneling_interaction(vec, theta, orbs, norb, nelec, spin)
        if gate == "num":
            return bg.apply_num_interaction(vec, theta, orbs[0], norb, nelec, spin)
        if gate == "num_num":
            return bg.apply_num_num_interaction(vec, theta, orbs, norb, nelec, spin)
        if gate == "hop":
            return bg.apply_hop_gate(vec, theta, orbs, norb, nelec, spin)
        if gate == "fsim":
            return bg.apply_fsim_gate(vec, theta, phi, orbs, norb, nelec, spin)
        if gate == "fswap":
            return bg.apply_fswap_gate(vec, orbs, norb, nelec, spin)
        if gate == "on_site":
            return bg.apply_on_site_interaction(vec, theta, orbs[0], norb, nelec)
        if gate == "num_op_prod":
            return bg.apply_num_op_prod_interaction(vec, theta, ([0, 2], [1]), norb, nelec)
        raise AssertionError(gate)

    def ucj_angles(self, vec, c, norb, nelec):
        from oracle import basic_gates as bg

        return bg.ucj_angles_apply(vec, norb, nelec, int(c["n_reps"]), c["params"], _pairs(c["pairs_aa"]),
                                   _pairs(c["pairs_ab"]), _pairs(c["givens_pairs"]), bool(c["final"]))

    def ucj_angles_from_ucj(self, vec, c, norb, nelec):
        # the angle form of a matrix-based operator applies the same unitary
        return self.m.ucj_spin_balanced_apply(vec, c["diag_coulomb_mats"], c["orbital_rotations"],
                                              opt(c["final_orbital_rotation"]), norb, nelec)


def _text(a):
    return bytes(np.asarray(a).astype(np.uint8)).decode()


def _pairs(a):
    return [tuple(int(x) for x in p) for p in np.asarray(a).reshape(-1, 2)]


class CudaAPI:
    """ffsim_b200: the public drop-in API over the CUDA library."""

    def __init__(self):
        import ffsim_b200 as ffsim

        self.f = ffsim
        self.apply_orbital_rotation = ffsim.apply_orbital_rotation
        self.apply_diag_coulomb_evolution = ffsim.apply_diag_coulomb_evolution
        self.apply_num_op_sum_evolution = ffsim.apply_num_op_sum_evolution
        self.contract_diag_coulomb = ffsim.contract_diag_coulomb
        self.contract_num_op_sum = ffsim.contract_num_op_sum

    def ucj(self, vec, c, norb, nelec):
        op = self.f.UCJOpSpinBalanced(c["diag_coulomb_mats"], c["orbital_rotations"],
                                      opt(c["final_orbital_rotation"]))
        return self.f.apply_unitary(vec, op, norb=norb, nelec=nelec)

    def trotter_df(self, vec, c, norb, nelec):
        ham = self.f.DoubleFactorizedHamiltonian(c["one_body_tensor"], c["diag_coulomb_mats"], c["orbital_rotations"],
                                                 constant=float(c["constant"]), z_representation=bool(c["z"]))
        return self.f.simulate_trotter_double_factorized(vec, ham, float(c["time"]), norb=norb, nelec=nelec,
                                                         n_steps=int(c["n_steps"]), order=int(c["order"]))

    def _dc(self, c):
        return self.f.DiagonalCoulombHamiltonian(c["one_body_tensor"], c["diag_coulomb_mats"], float(c["constant"]))

    def dc_matvec(self, vec, c, norb, nelec):
        return self.f.linear_operator(self._dc(c), norb=norb, nelec=nelec) @ vec

    def dc_split_op(self, vec, c, norb, nelec):
        return self.f.simulate_trotter_diag_coulomb_split_op(vec, self._dc(c), float(c["time"]), norb=norb, nelec=nelec,
                                                             n_steps=int(c["n_steps"]), order=int(c["order"]))

    def ucj_unbalanced(self, vec, c, norb, nelec):
        op = self.f.UCJOpSpinUnbalanced(c["diag_coulomb_mats"], c["orbital_rotations"],
                                        opt(c["final_orbital_rotation"]))
        return self.f.apply_unitary(vec, op, norb=norb, nelec=nelec)

    def ucj_spinless(self, vec, c, norb, nelec):
        op = self.f.UCJOpSpinless(c["diag_coulomb_mats"], c["orbital_rotations"], opt(c["final_orbital_rotation"]))
        return self.f.apply_unitary(vec, op, norb=norb, nelec=nelec)

    def _df(self, c):
        return self.f.DoubleFactorizedHamiltonian(c["one_body_tensor"], c["diag_coulomb_mats"], c["orbital_rotations"],
                                                  constant=float(c["constant"]), z_representation=bool(c["z"]))

    def df_matvec(self, vec, c, norb, nelec):
        return self.f.linear_operator(self._df(c), norb=norb, nelec=nelec) @ vec

    def qdrift(self, vec, c, norb, nelec):
        return self.f.simulate_qdrift_double_factorized(
            vec, self._df(c), float(c["time"]), norb=norb, nelec=nelec, n_steps=int(c["n_steps"]),
            symmetric=bool(c["symmetric"]), probabilities=bytes(c["probabilities"].astype(np.uint8)).decode(),
            n_samples=int(c["n_samples"]), seed=int(c["seed"]))


def _cuda_basic(f, vec, c, norb, nelec):
    gate, sname = _text(c["gate"]), _text(c["spin"])
    spin = {"a": f.Spin.ALPHA, "b": f.Spin.BETA, "ab": f.Spin.ALPHA_AND_BETA}[sname]
    theta, phi, orbs = float(c["theta"]), float(c["phi"]), tuple(int(x) for x in c["orbs"])
    kw = dict(norb=norb, nelec=nelec, spin=spin)
    if gate == "givens":
        return f.apply_givens_rotation(vec, theta, orbs, phi=phi, **kw)
    if gate == "tunneling":
        return f.apply_tunneling_interaction(vec, theta, orbs, **kw)
    if gate == "num":
        return f.apply_num_interaction(vec, theta, orbs[0], **kw)
    if gate == "num_num":
        return f.apply_num_num_interaction(vec, theta, orbs, **kw)
    if gate == "hop":
        return f.apply_hop_gate(vec, theta, orbs, **kw)
    if gate == "fsim":
        return f.apply_fsim_gate(vec, theta, phi, orbs, **kw)
    if gate == "fswap":
        return f.apply_fswap_gate(vec, orbs, **kw)
    if gate == "on_site":
        return f.apply_on_site_interaction(vec, theta, orbs[0], norb=norb, nelec=nelec)
    if gate == "num_op_prod":
        return f.apply_num_op_prod_interaction(vec, theta, ([0, 2], [1]), norb=norb, nelec=nelec)
    raise AssertionError(gate)


def _cuda_ucj_angles(f, vec, c, norb, nelec):
    op = f.UCJAnglesOpSpinBalanced.from_parameters(
        c["params"], norb=norb, n_reps=int(c["n_reps"]),
        num_num_interaction_pairs=(_pairs(c["pairs_aa"]), _pairs(c["pairs_ab"])),
        givens_interaction_pairs=_pairs(c["givens_pairs"]), with_final_givens_ansatz_op=bool(c["final"]))
    assert np.array_equal(op.to_parameters(), c["roundtrip"])
    return f.apply_unitary(vec, op, norb=norb, nelec=nelec)


def _cuda_ucj_angles_from_ucj(f, vec, c, norb, nelec):
    ucj = f.UCJOpSpinBalanced(c["diag_coulomb_mats"], c["orbital_rotations"], opt(c["final_orbital_rotation"]))
    op = f.UCJAnglesOpSpinBalanced.from_ucj_op(ucj)
    assert np.allclose(op.to_parameters(), c["params"], rtol=0, atol=1e-12)
    return f.apply_unitary(vec, op, norb=norb, nelec=nelec)


CudaAPI.basic = lambda self, vec, c, norb, nelec: _cuda_basic(self.f, vec, c, norb, nelec)
CudaAPI.ucj_angles = lambda self, vec, c, norb, nelec: _cuda_ucj_angles(self.f, vec, c, norb, nelec)
CudaAPI.ucj_angles_from_ucj = lambda self, vec, c, norb, nelec: _cuda_ucj_angles_from_ucj(self.f, vec, c, norb, nelec)


def run_case(api, c):
    kind = str(c["kind"])
    norb, nelec, vec = int(c["norb"]), nelec_of(c), c["vec"]
    before = vec.copy()
    if kind == "orbital_rotation":
        ma, mb = opt(c["mat_a"]), opt(c["mat_b"])
        mat = ma if (ma is not None and mb is not None and np.array_equal(ma, mb)) else (ma, mb)
        got = api.apply_orbital_rotation(vec, mat, norb, nelec)
    elif kind == "orbital_rotation_spinless":
        got = api.apply_orbital_rotation(vec, c["mat_a"], norb, nelec)
    elif kind == "diag_coulomb":
        mat = c["mat_aa"] if int(c["single"]) else (opt(c["mat_aa"]), opt(c["mat_ab"]), opt(c["mat_bb"]))
        got = api.apply_diag_coulomb_evolution(vec, mat, float(c["time"]), norb, nelec, orbital_rotation=opt(c["rot"]),
                                               z_representation=bool(c["z"]))
    elif kind == "diag_coulomb_spinless":
        got = api.apply_diag_coulomb_evolution(vec, c["mat_aa"], float(c["time"]), norb, nelec)
    elif kind == "num_op_sum":
        ca, cb = opt(c["coeffs_a"]), opt(c["coeffs_b"])
        coeffs = ca if (ca is not None and cb is not None and np.array_equal(ca, cb)) else (ca, cb)
        got = api.apply_num_op_sum_evolution(vec, coeffs, float(c["time"]), norb, nelec, orbital_rotation=opt(c["rot"]))
    elif kind == "contract_diag_coulomb":
        mat = c["mat_aa"] if int(c["single"]) else (c["mat_aa"], c["mat_ab"], c["mat_bb"])
        got = api.contract_diag_coulomb(vec, mat, norb, nelec, z_representation=bool(c["z"]))
    elif kind == "contract_num_op_sum":
        got = api.contract_num_op_sum(vec, c["coeffs_a"], norb, nelec)
    elif kind == "ucj":
        got = api.ucj(vec, c, norb, nelec)
    elif kind == "trotter_df":
        got = api.trotter_df(vec, c, norb, nelec)
    elif kind == "dc_matvec":
        got = api.dc_matvec(vec, c, norb, nelec)
    elif kind == "dc_split_op":
        got = api.dc_split_op(vec, c, norb, nelec)
    elif kind in ("ucj_unbalanced", "ucj_spinless", "df_matvec", "qdrift", "basic", "ucj_angles", "ucj_angles_from_ucj"):
        got = getattr(api, kind)(vec, c, norb, nelec)
    else:
        raise AssertionError(kind)
    assert np.array_equal(vec, before), "the input vector was modified (copy=True semantics)"
    return got


STATE_CASES = [n for n in CASES if "vec" in CASES[n] and "expected" in CASES[n]]


def test_fixture_is_complete():
    kinds = {str(c["kind"]) for c in CASES.values()}
    assert kinds >= {"orbital_rotation", "orbital_rotation_spinless", "diag_coulomb", "diag_coulomb_spinless",
                     "num_op_sum", "contract_diag_coulomb", "contract_num_op_sum", "ucj", "trotter_df",
                     "dc_matvec", "dc_split_op", "zero_one", "one", "random_unitary",
                     "ucj_unbalanced", "ucj_spinless", "df_matvec", "qdrift",
                     "basic", "ucj_angles", "ucj_angles_from_ucj", "params_balanced", "params_balanced_from",
                     "params_unbalanced", "params_spinless", "givens_from_rotation", "qdrift_probs"}
    assert len(STATE_CASES) >= 150


# ----------------------------------------------------------------------------- CPU: pin the oracle
@pytest.mark.parametrize("name", STATE_CASES)
def test_oracle_reproduces_reference(name):
    c = CASES[name]
    err = rel_err(run_case(OracleAPI(), c), c["expected"])
    assert err <= TOL, f"{name}: rel 2-norm error {err:.3e}"


@pytest.mark.parametrize("name", names("random/"))
def test_generators_bit_exact(name):
    """oracle/rand.py and ffsim_b200/random.py against python/ffsim/random/random.py."""
    import ffsim_b200.random as prod
    from oracle import rand

    c = CASES[name]
    kind, n, seed = str(c["kind"]), int(c["n"]), int(c["seed"])
    for mod in (rand, prod):
        got = getattr(mod, kind)(n, seed=seed)
        assert np.array_equal(got, c["expected"]), f"{mod.__name__}.{kind}"


@pytest.mark.parametrize("name", names("random_op/"))
def test_operator_generators_bit_exact(name):
    """ffsim_b200/random.py against python/ffsim/random/random.py:668-880 (UCJ spin-unbalanced, spinless)."""
    import ffsim_b200.random as prod

    c = CASES[name]
    op = getattr(prod, str(c["kind"]))(int(c["n"]), n_reps=int(c["n_reps"]),
                                       with_final_orbital_rotation=bool(c["final"]),
                                       diag_coulomb_normal=bool(c["normal"]), diag_coulomb_mean=float(c["mean"]),
                                       seed=int(c["seed"]))
    assert np.array_equal(op.diag_coulomb_mats, c["diag_coulomb_mats"])
    assert np.array_equal(op.orbital_rotations, c["orbital_rotations"])
    want_final = opt(c["final_orbital_rotation"])
    assert (op.final_orbital_rotation is None) == (want_final is None)
    if want_final is not None:
        assert np.array_equal(op.final_orbital_rotation, want_final)


@pytest.mark.parametrize("name", names("tables/"))
def test_tables_bit_exact_vs_reference_argsort(name):
    """The reference's argsort construction (gates/orbital_rotation.py:203-236) against the
    direct construction of oracle/cistring.py and of the C ABI (csrc/tables.cpp)."""
    from ffsim_b200 import cistring as prod
    from oracle import cistring

    c = CASES[name]
    norb, nocc, i = int(c["norb"]), int(c["nocc"]), int(c["i"])
    if str(c["kind"]) == "zero_one":
        j = int(c["j"])
        want = c["expected"]
        assert np.array_equal(cistring.zero_one_subspace_indices(norb, nocc, (i, j)).astype(np.int64), want)
        assert np.array_equal(np.asarray(prod.zero_one_subspace_indices(norb, nocc, (i, j))).astype(np.int64), want)
    else:
        want = c["expected"]
        assert np.array_equal(cistring.one_subspace_indices(norb, nocc, (i,)).astype(np.int64), want)
        assert np.array_equal(np.asarray(prod.one_subspace_indices(norb, nocc, (i,))).astype(np.int64), want)


# ----------------------------------------------------------------------------- CPU: host-side product code
@pytest.mark.parametrize("name", names("params/"))
def test_parameter_vectors_match_reference(name):
    """n_params / to_parameters / from_parameters of the UCJ operators (the reference's vector layout)."""
    import ffsim_b200 as f

    c = CASES[name]
    kind, norb, n_reps, final = str(c["kind"]), int(c["norb"]), int(c["n_reps"]), bool(c["final"])
    if kind in ("params_balanced", "params_balanced_from"):
        pairs = None if opt(c["pairs_aa"]) is None else (_pairs(c["pairs_aa"]), _pairs(c["pairs_ab"]))
        if kind == "params_balanced":
            op = f.UCJOpSpinBalanced(c["diag_coulomb_mats"], c["orbital_rotations"], c["final_orbital_rotation"])
            got = op.to_parameters(interaction_pairs=pairs)
            assert got.shape == c["expected"].shape and np.allclose(got, c["expected"], rtol=0, atol=1e-12)
            assert len(got) == f.UCJOpSpinBalanced.n_params(norb, n_reps, interaction_pairs=pairs,
                                                             with_final_orbital_rotation=final)
            back = f.UCJOpSpinBalanced.from_parameters(got, norb=norb, n_reps=n_reps, interaction_pairs=pairs,
                                                       with_final_orbital_rotation=final)
            assert np.allclose(back.orbital_rotations, op.orbital_rotations, atol=1e-12)
        else:
            op = f.UCJOpSpinBalanced.from_parameters(c["params"], norb=norb, n_reps=n_reps, interaction_pairs=pairs,
                                                     with_final_orbital_rotation=final)
            assert np.allclose(op.diag_coulomb_mats, c["diag_coulomb_mats"], rtol=0, atol=1e-13)
            assert np.allclose(op.orbital_rotations, c["orbital_rotations"], rtol=0, atol=1e-12)
            assert np.allclose(op.final_orbital_rotation, c["final_orbital_rotation"], rtol=0, atol=1e-12)
    elif kind == "params_unbalanced":
        op = f.UCJOpSpinUnbalanced(c["diag_coulomb_mats"], c["orbital_rotations"], c["final_orbital_rotation"])
        got = op.to_parameters()
        assert len(got) == int(c["n_params"]) == f.UCJOpSpinUnbalanced.n_params(norb, n_reps,
                                                                                  with_final_orbital_rotation=final)
        assert np.allclose(got, c["expected"], rtol=0, atol=1e-12)
        back = f.UCJOpSpinUnbalanced.from_parameters(got, norb=norb, n_reps=n_reps, with_final_orbital_rotation=final)
        assert np.allclose(back.diag_coulomb_mats, op.diag_coulomb_mats, atol=1e-13)
        assert np.allclose(back.final_orbital_rotation, op.final_orbital_rotation, atol=1e-12)
    else:
        pairs = _pairs(c["pairs"])
        op = f.UCJOpSpinless(c["diag_coulomb_mats"], c["orbital_rotations"])
        got = op.to_parameters(interaction_pairs=pairs)
        assert len(got) == int(c["n_params"]) == f.UCJOpSpinless.n_params(norb, n_reps, interaction_pairs=pairs)
        assert np.allclose(got, c["expected"], rtol=0, atol=1e-12)


def test_parameter_vector_errors():
    import ffsim_b200 as f

    with pytest.raises(ValueError, match="Expected 36 but got 3"):
        f.UCJOpSpinBalanced.from_parameters(np.zeros(3), norb=4, n_reps=1)
    with pytest.raises(ValueError, match="Duplicate interaction pairs"):
        f.UCJOpSpinBalanced.n_params(4, 1, interaction_pairs=([(0, 1), (0, 1)], None))
    with pytest.raises(ValueError, match="lower triangular pair"):
        f.UCJOpSpinless.n_params(4, 1, interaction_pairs=[(2, 1)])


def test_givens_ansatz_from_orbital_rotation_matches_reference():
    """GivensAnsatzOp.from_orbital_rotation: brickwork layout, angles and the rebuilt unitary."""
    import ffsim_b200 as f

    c = CASES["angles/givens_from_rotation_5"]
    g = f.GivensAnsatzOp.from_orbital_rotation(c["mat"])
    assert [tuple(p) for p in g.interaction_pairs] == _pairs(c["pairs"])
    assert np.allclose(g.thetas, c["thetas"], rtol=0, atol=1e-12)
    assert np.allclose(g.phis, c["phis"], rtol=0, atol=1e-12)
    assert np.allclose(g.phase_angles, c["phase_angles"], rtol=0, atol=1e-12)
    assert np.allclose(g.to_orbital_rotation(), c["rotation"], rtol=0, atol=1e-13)
    assert np.allclose(g.to_orbital_rotation(), c["mat"], rtol=0, atol=1e-12)
    back = f.GivensAnsatzOp.from_parameters(g.to_parameters(), norb=5, interaction_pairs=g.interaction_pairs)
    assert back._approx_eq_(g, 0, 1e-15)


@pytest.mark.parametrize("name", names("qdrift_probs/"))
def test_qdrift_probabilities_match_reference(name):
    """qdrift_probabilities incl. the Slater-determinant-optimal weights (trotter/qdrift.py:247-348, states/wick.py)."""
    import ffsim_b200 as f

    c = CASES[name]
    ham = f.DoubleFactorizedHamiltonian(c["one_body_tensor"], c["diag_coulomb_mats"], c["orbital_rotations"],
                                        constant=float(c["constant"]), z_representation=bool(c["z"]))
    got = f.qdrift_probabilities(ham, sampling_method=_text(c["method"]), nelec=nelec_of(c), one_rdm=c["one_rdm"])
    assert np.allclose(got, c["expected"], rtol=1e-10, atol=1e-13), (got, c["expected"])


# ----------------------------------------------------------------------------- GPU: the product
@pytest.mark.gpu
@pytest.mark.parametrize("name", STATE_CASES)
def test_cuda_reproduces_reference(name):
    c = CASES[name]
    err = rel_err(run_case(CudaAPI(), c), c["expected"])
    assert err <= TOL, f"{name}: rel 2-norm error {err:.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["orbital_rotation/pair_9_4_3", "ucj/lucj_8_4_4_L2", "diag_coulomb/rotated_z_6_2_3"])
def test_cuda_device_tensor_reproduces_reference(name):
    """Same calls with a CUDA tensor in / CUDA tensor out."""
    import torch

    c = dict(CASES[name])
    host = c["vec"]

    class DevAPI(CudaAPI):
        pass

    api = DevAPI()
    dev = torch.from_numpy(host).cuda()
    c["vec"] = host  # run_case checks copy semantics on the host array; run the device call separately
    kind = str(c["kind"])
    norb, nelec = int(c["norb"]), nelec_of(c)
    if kind == "orbital_rotation":
        got = api.apply_orbital_rotation(dev, (c["mat_a"], c["mat_b"]), norb, nelec)
    elif kind == "ucj":
        got = api.ucj(dev, c, norb, nelec)
    else:
        got = api.apply_diag_coulomb_evolution(dev, (c["mat_aa"], c["mat_ab"], c["mat_bb"]), float(c["time"]), norb,
                                               nelec, orbital_rotation=c["rot"], z_representation=bool(c["z"]))
    assert isinstance(got, torch.Tensor) and got.is_cuda
    assert torch.equal(dev.cpu(), torch.from_numpy(host)), "copy=True must not modify the device input"
    assert rel_err(got.cpu().numpy(), c["expected"]) <= TOL
