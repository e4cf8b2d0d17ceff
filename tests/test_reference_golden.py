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
    elif kind in ("ucj_unbalanced", "ucj_spinless", "df_matvec", "qdrift"):
        got = getattr(api, kind)(vec, c, norb, nelec)
    else:
        raise AssertionError(kind)
    assert np.array_equal(vec, before), "the input vector was modified (copy=True semantics)"
    return got


STATE_CASES = [n for n in CASES if not n.startswith(("random/", "random_op/", "tables/"))]


def test_fixture_is_complete():
    kinds = {str(c["kind"]) for c in CASES.values()}
    assert kinds >= {"orbital_rotation", "orbital_rotation_spinless", "diag_coulomb", "diag_coulomb_spinless",
                     "num_op_sum", "contract_diag_coulomb", "contract_num_op_sum", "ucj", "trotter_df",
                     "dc_matvec", "dc_split_op", "zero_one", "one", "random_unitary",
                     "ucj_unbalanced", "ucj_spinless", "df_matvec", "qdrift"}
    assert len(STATE_CASES) >= 93


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
