"""Generate tests/golden/ref_*.npz by running the REFERENCE's own Python code (see ref_shim.py).

    python tests/golden/make_golden.py          # in the build container only

Every case is produced by the reference's public drivers (python/ffsim/gates/*.py,
contract/*.py, variational/ucj_spin_balanced.py, trotter/*.py, hamiltonians/*.py) with
inputs from the reference's own generators (python/ffsim/random/random.py), executing the
reference's pure-Python kernel twins (python/ffsim/_slow/**).  Inputs and outputs are both
stored, so the tests do not depend on any generator of this repository.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

import ffsim.random.random as rr  # noqa: E402
from ffsim.contract.diag_coulomb import contract_diag_coulomb  # noqa: E402
from ffsim.contract.num_op_sum import contract_num_op_sum  # noqa: E402
from ffsim.gates.diag_coulomb import apply_diag_coulomb_evolution  # noqa: E402
from ffsim.gates.num_op_sum import apply_num_op_sum_evolution  # noqa: E402
from ffsim.gates.orbital_rotation import (  # noqa: E402
    _one_subspace_indices,
    _zero_one_subspace_indices,
    apply_orbital_rotation,
)
from ffsim.hamiltonians.diagonal_coulomb_hamiltonian import DiagonalCoulombHamiltonian  # noqa: E402
from ffsim.protocols.apply_unitary_protocol import apply_unitary  # noqa: E402
from ffsim.protocols.linear_operator_protocol import linear_operator  # noqa: E402
from ffsim.states.dimensions import dim as ref_dim  # noqa: E402
from ffsim.states.slater import hartree_fock_state  # noqa: E402
from ffsim.trotter.diagonal_coulomb_split_op import simulate_trotter_diag_coulomb_split_op  # noqa: E402
from ffsim.trotter.double_factorized import simulate_trotter_double_factorized  # noqa: E402
from ffsim.trotter.qdrift import simulate_qdrift_double_factorized  # noqa: E402

CASES: dict[str, dict] = {}
NONE = np.zeros((0,))  # stands for a `None` member


def opt(x):
    return NONE if x is None else np.asarray(x)


def add(name, **arrays):
    assert name not in CASES, name
    CASES[name] = {k: np.asarray(v) for k, v in arrays.items()}


def main():
    # ---- generators (pins oracle/rand.py and ffsim_b200/random.py) ----
    add("random/unitary_5_seed11", kind="random_unitary", n=5, seed=11, expected=rr.random_unitary(5, seed=11))
    add("random/real_symmetric_6_seed12", kind="random_real_symmetric_matrix", n=6, seed=12,
        expected=rr.random_real_symmetric_matrix(6, seed=12))
    add("random/state_vector_20_seed13", kind="random_state_vector", n=20, seed=13,
        expected=rr.random_state_vector(20, seed=13))

    # ---- address tables (reference argsort construction, orbital_rotation.py:203-236) ----
    for norb, nocc in [(4, 2), (6, 3), (7, 2), (8, 5)]:
        for (i, j) in sorted({(0, 1), (1, 0), (norb - 2, norb - 1), (2, 3), (3, 2)}):
            add(f"tables/zero_one_{norb}_{nocc}_{i}_{j}", kind="zero_one", norb=norb, nocc=nocc, i=i, j=j,
                expected=_zero_one_subspace_indices(norb, nocc, (i, j)).astype(np.int64))
        add(f"tables/one_{norb}_{nocc}", kind="one", norb=norb, nocc=nocc, i=1,
            expected=_one_subspace_indices(norb, nocc, (1,)).astype(np.int64))

    # ---- orbital rotation ----
    rng = np.random.default_rng(20261017)
    for norb, nelec in [(4, (2, 2)), (5, (3, 2)), (6, (3, 3)), (6, (1, 4)), (3, (0, 2)), (7, (3, 2))]:
        d = ref_dim(norb, nelec)
        vec = rr.random_state_vector(d, seed=rng)
        ua, ub = rr.random_unitary(norb, seed=rng), rr.random_unitary(norb, seed=rng)
        tag = f"{norb}_{nelec[0]}_{nelec[1]}"
        add(f"orbital_rotation/same_{tag}", kind="orbital_rotation", norb=norb, nelec=nelec, vec=vec,
            mat_a=ua, mat_b=ua, expected=apply_orbital_rotation(vec, ua, norb, nelec))
        add(f"orbital_rotation/pair_{tag}", kind="orbital_rotation", norb=norb, nelec=nelec, vec=vec,
            mat_a=ua, mat_b=ub, expected=apply_orbital_rotation(vec, (ua, ub), norb, nelec))
        add(f"orbital_rotation/alpha_only_{tag}", kind="orbital_rotation", norb=norb, nelec=nelec, vec=vec,
            mat_a=ua, mat_b=NONE, expected=apply_orbital_rotation(vec, (ua, None), norb, nelec))
        add(f"orbital_rotation/beta_only_{tag}", kind="orbital_rotation", norb=norb, nelec=nelec, vec=vec,
            mat_a=NONE, mat_b=ub, expected=apply_orbital_rotation(vec, (None, ub), norb, nelec))
    for norb, nocc in [(5, 3), (6, 2)]:
        d = ref_dim(norb, nocc)
        vec = rr.random_state_vector(d, seed=rng)
        u = rr.random_unitary(norb, seed=rng)
        add(f"orbital_rotation/spinless_{norb}_{nocc}", kind="orbital_rotation_spinless", norb=norb, nelec=nocc,
            vec=vec, mat_a=u, expected=apply_orbital_rotation(vec, u, norb, nocc))
    # one mid-size case: several tiles / register-block classes on the CUDA side
    norb, nelec = 9, (4, 3)
    vec = rr.random_state_vector(ref_dim(norb, nelec), seed=rng)
    ua, ub = rr.random_unitary(norb, seed=rng), rr.random_unitary(norb, seed=rng)
    add("orbital_rotation/pair_9_4_3", kind="orbital_rotation", norb=norb, nelec=nelec, vec=vec,
        mat_a=ua, mat_b=ub, expected=apply_orbital_rotation(vec, (ua, ub), norb, nelec))
    # a permutation matrix and the identity: rotations that degenerate (c = 0 / no rotations)
    norb, nelec = 5, (2, 3)
    vec = rr.random_state_vector(ref_dim(norb, nelec), seed=rng)
    perm = np.eye(norb)[:, [2, 0, 4, 1, 3]].astype(complex)
    add("orbital_rotation/permutation_5_2_3", kind="orbital_rotation", norb=norb, nelec=nelec, vec=vec,
        mat_a=perm, mat_b=perm, expected=apply_orbital_rotation(vec, perm, norb, nelec))
    add("orbital_rotation/identity_5_2_3", kind="orbital_rotation", norb=norb, nelec=nelec, vec=vec,
        mat_a=np.eye(norb, dtype=complex), mat_b=np.eye(norb, dtype=complex),
        expected=apply_orbital_rotation(vec, np.eye(norb, dtype=complex), norb, nelec))

    # ---- diagonal Coulomb evolution ----
    for norb, nelec in [(4, (2, 2)), (5, (3, 2)), (6, (2, 3))]:
        d = ref_dim(norb, nelec)
        tag = f"{norb}_{nelec[0]}_{nelec[1]}"
        vec = rr.random_state_vector(d, seed=rng)
        u = rr.random_unitary(norb, seed=rng)
        maa = rr.random_real_symmetric_matrix(norb, seed=rng)
        mab = rng.standard_normal((norb, norb))  # alpha-beta block need not be symmetric
        mbb = rr.random_real_symmetric_matrix(norb, seed=rng)
        t = 0.73
        for z in (False, True):
            zt = "z" if z else "num"
            add(f"diag_coulomb/sym_{zt}_{tag}", kind="diag_coulomb", norb=norb, nelec=nelec, vec=vec, time=t, z=z,
                mat_aa=maa, mat_ab=maa, mat_bb=maa, rot=NONE, single=1,
                expected=apply_diag_coulomb_evolution(vec, maa, t, norb, nelec, z_representation=z))
            add(f"diag_coulomb/triple_{zt}_{tag}", kind="diag_coulomb", norb=norb, nelec=nelec, vec=vec, time=t, z=z,
                mat_aa=maa, mat_ab=mab, mat_bb=mbb, rot=NONE, single=0,
                expected=apply_diag_coulomb_evolution(vec, (maa, mab, mbb), t, norb, nelec, z_representation=z))
            add(f"diag_coulomb/ab_only_{zt}_{tag}", kind="diag_coulomb", norb=norb, nelec=nelec, vec=vec, time=t, z=z,
                mat_aa=NONE, mat_ab=mab, mat_bb=NONE, rot=NONE, single=0,
                expected=apply_diag_coulomb_evolution(vec, (None, mab, None), t, norb, nelec, z_representation=z))
            add(f"diag_coulomb/rotated_{zt}_{tag}", kind="diag_coulomb", norb=norb, nelec=nelec, vec=vec, time=t, z=z,
                mat_aa=maa, mat_ab=mab, mat_bb=mbb, rot=u, single=0,
                expected=apply_diag_coulomb_evolution(vec, (maa, mab, mbb), t, norb, nelec, orbital_rotation=u,
                                                      z_representation=z))
    norb, nocc = 5, 2
    vec = rr.random_state_vector(ref_dim(norb, nocc), seed=rng)
    m = rr.random_real_symmetric_matrix(norb, seed=rng)
    add("diag_coulomb/spinless_5_2", kind="diag_coulomb_spinless", norb=norb, nelec=nocc, vec=vec, time=0.4,
        mat_aa=m, expected=apply_diag_coulomb_evolution(vec, m, 0.4, norb, nocc))

    # ---- number-operator-sum evolution ----
    for norb, nelec in [(4, (2, 2)), (6, (3, 2))]:
        d = ref_dim(norb, nelec)
        tag = f"{norb}_{nelec[0]}_{nelec[1]}"
        vec = rr.random_state_vector(d, seed=rng)
        u = rr.random_unitary(norb, seed=rng)
        ca, cb = rng.standard_normal(norb), rng.standard_normal(norb)
        t = 1.3
        add(f"num_op_sum/same_{tag}", kind="num_op_sum", norb=norb, nelec=nelec, vec=vec, time=t,
            coeffs_a=ca, coeffs_b=ca, rot=NONE, expected=apply_num_op_sum_evolution(vec, ca, t, norb, nelec))
        add(f"num_op_sum/pair_{tag}", kind="num_op_sum", norb=norb, nelec=nelec, vec=vec, time=t,
            coeffs_a=ca, coeffs_b=cb, rot=NONE, expected=apply_num_op_sum_evolution(vec, (ca, cb), t, norb, nelec))
        add(f"num_op_sum/beta_only_{tag}", kind="num_op_sum", norb=norb, nelec=nelec, vec=vec, time=t,
            coeffs_a=NONE, coeffs_b=cb, rot=NONE,
            expected=apply_num_op_sum_evolution(vec, (None, cb), t, norb, nelec))
        add(f"num_op_sum/rotated_{tag}", kind="num_op_sum", norb=norb, nelec=nelec, vec=vec, time=t,
            coeffs_a=ca, coeffs_b=cb, rot=u,
            expected=apply_num_op_sum_evolution(vec, (ca, cb), t, norb, nelec, orbital_rotation=u))

    # ---- contractions ----
    for norb, nelec in [(4, (2, 2)), (5, (2, 3))]:
        d = ref_dim(norb, nelec)
        tag = f"{norb}_{nelec[0]}_{nelec[1]}"
        vec = rr.random_state_vector(d, seed=rng)
        maa = rr.random_real_symmetric_matrix(norb, seed=rng)
        mab = rng.standard_normal((norb, norb))
        mbb = rr.random_real_symmetric_matrix(norb, seed=rng)
        ca, cb = rng.standard_normal(norb), rng.standard_normal(norb)
        for z in (False, True):
            zt = "z" if z else "num"
            add(f"contract_diag_coulomb/sym_{zt}_{tag}", kind="contract_diag_coulomb", norb=norb, nelec=nelec,
                vec=vec, z=z, mat_aa=maa, mat_ab=maa, mat_bb=maa, single=1,
                expected=contract_diag_coulomb(vec, maa, norb, nelec, z_representation=z))
            add(f"contract_diag_coulomb/triple_{zt}_{tag}", kind="contract_diag_coulomb", norb=norb, nelec=nelec,
                vec=vec, z=z, mat_aa=maa, mat_ab=mab, mat_bb=mbb, single=0,
                expected=contract_diag_coulomb(vec, (maa, mab, mbb), norb, nelec, z_representation=z))
        add(f"contract_num_op_sum/same_{tag}", kind="contract_num_op_sum", norb=norb, nelec=nelec, vec=vec,
            coeffs_a=ca, coeffs_b=ca, expected=contract_num_op_sum(vec, ca, norb, nelec))

    # ---- UCJOpSpinBalanced / LUCJ (variational/ucj_spin_balanced.py:657) ----
    for norb, nelec, n_reps, final, lucj in [(4, (2, 2), 2, True, True), (5, (3, 2), 3, False, False),
                                            (6, (3, 3), 2, True, True), (8, (4, 4), 2, True, True)]:
        pairs = None
        if lucj:  # docs/explanations/lucj.ipynb:392-393
            pairs = ([(p, p + 1) for p in range(norb - 1)], [(p, p) for p in range(norb)])
        op = rr.random_ucj_op_spin_balanced(norb, n_reps=n_reps, interaction_pairs=pairs,
                                            with_final_orbital_rotation=final, seed=rng)
        vec = hartree_fock_state(norb, nelec)
        tag = f"{norb}_{nelec[0]}_{nelec[1]}_L{n_reps}"
        add(f"ucj/{'lucj' if lucj else 'full'}_{tag}", kind="ucj", norb=norb, nelec=nelec, vec=vec,
            diag_coulomb_mats=op.diag_coulomb_mats, orbital_rotations=op.orbital_rotations,
            final_orbital_rotation=opt(op.final_orbital_rotation),
            expected=apply_unitary(vec, op, norb=norb, nelec=nelec))

    # ---- double-factorized Trotter (trotter/double_factorized.py:25) ----
    for norb, nelec, rank, z, order, n_steps in [(4, (2, 2), 3, False, 0, 1), (4, (2, 2), 3, False, 1, 2),
                                                 (5, (2, 3), 4, True, 0, 2), (4, (1, 2), 2, False, 2, 1)]:
        ham = rr.random_double_factorized_hamiltonian(norb, rank=rank, z_representation=z, seed=rng)
        vec = rr.random_state_vector(ref_dim(norb, nelec), seed=rng)
        t = 0.35
        tag = f"{norb}_{nelec[0]}_{nelec[1]}_r{rank}_{'z' if z else 'num'}_o{order}_s{n_steps}"
        add(f"trotter_df/{tag}", kind="trotter_df", norb=norb, nelec=nelec, vec=vec, time=t, order=order,
            n_steps=n_steps, z=z, one_body_tensor=ham.one_body_tensor, diag_coulomb_mats=ham.diag_coulomb_mats,
            orbital_rotations=ham.orbital_rotations, constant=ham.constant,
            expected=simulate_trotter_double_factorized(vec, ham, t, norb=norb, nelec=nelec, n_steps=n_steps,
                                                        order=order))

    # ---- DiagonalCoulombHamiltonian: matvec and split-operator Trotter ----
    for norb, nelec in [(4, (2, 2)), (5, (3, 2))]:
        ham = rr.random_diagonal_coulomb_hamiltonian(norb, seed=rng)
        vec = rr.random_state_vector(ref_dim(norb, nelec), seed=rng)
        tag = f"{norb}_{nelec[0]}_{nelec[1]}"
        lin = linear_operator(ham, norb=norb, nelec=nelec)
        add(f"dc_hamiltonian/matvec_{tag}", kind="dc_matvec", norb=norb, nelec=nelec, vec=vec,
            one_body_tensor=ham.one_body_tensor, diag_coulomb_mats=ham.diag_coulomb_mats, constant=ham.constant,
            expected=lin @ vec)
        add(f"dc_hamiltonian/split_op_{tag}", kind="dc_split_op", norb=norb, nelec=nelec, vec=vec, time=0.2,
            order=1, n_steps=2, one_body_tensor=ham.one_body_tensor, diag_coulomb_mats=ham.diag_coulomb_mats,
            constant=ham.constant,
            expected=simulate_trotter_diag_coulomb_split_op(vec, ham, 0.2, norb=norb, nelec=nelec, n_steps=2, order=1))

    # ---- SURVEY.md section 8f rows (appended: the cases above keep their random streams) ----
    # UCJOpSpinUnbalanced (variational/ucj_spin_unbalanced.py:705): per-spin rotations, J_ab not symmetric
    for norb, nelec, n_reps, final in [(4, (2, 2), 2, True), (5, (3, 2), 2, False), (6, (2, 4), 1, True)]:
        op = rr.random_ucj_op_spin_unbalanced(norb, n_reps=n_reps, with_final_orbital_rotation=final, seed=rng)
        vec = rr.random_state_vector(ref_dim(norb, nelec), seed=rng)
        tag = f"{norb}_{nelec[0]}_{nelec[1]}_L{n_reps}"
        add(f"ucj_unbalanced/{tag}", kind="ucj_unbalanced", norb=norb, nelec=nelec, vec=vec,
            diag_coulomb_mats=op.diag_coulomb_mats, orbital_rotations=op.orbital_rotations,
            final_orbital_rotation=opt(op.final_orbital_rotation),
            expected=apply_unitary(vec, op, norb=norb, nelec=nelec))
    # UCJOpSpinless (variational/ucj_spinless.py:456): integer nelec and pair nelec
    for norb, nelec, n_reps, final in [(5, 2, 2, True), (6, 3, 1, False), (4, (2, 1), 2, True), (5, (2, 2), 1, False)]:
        op = rr.random_ucj_op_spinless(norb, n_reps=n_reps, with_final_orbital_rotation=final, seed=rng)
        vec = rr.random_state_vector(ref_dim(norb, nelec), seed=rng)
        tag = f"{norb}_{nelec}_L{n_reps}" if isinstance(nelec, int) else f"{norb}_{nelec[0]}_{nelec[1]}_L{n_reps}"
        add(f"ucj_spinless/{tag}", kind="ucj_spinless", norb=norb, nelec=nelec, vec=vec,
            diag_coulomb_mats=op.diag_coulomb_mats, orbital_rotations=op.orbital_rotations,
            final_orbital_rotation=opt(op.final_orbital_rotation),
            expected=apply_unitary(vec, op, norb=norb, nelec=nelec))
    # DoubleFactorizedHamiltonian._linear_operator_ (hamiltonians/double_factorized_hamiltonian.py:244)
    for norb, nelec, rank, z in [(4, (2, 2), 3, False), (5, (3, 2), 4, True)]:
        ham = rr.random_double_factorized_hamiltonian(norb, rank=rank, z_representation=z, seed=rng)
        vec = rr.random_state_vector(ref_dim(norb, nelec), seed=rng)
        tag = f"{norb}_{nelec[0]}_{nelec[1]}_r{rank}_{'z' if z else 'num'}"
        add(f"df_hamiltonian/matvec_{tag}", kind="df_matvec", norb=norb, nelec=nelec, vec=vec, z=z,
            one_body_tensor=ham.one_body_tensor, diag_coulomb_mats=ham.diag_coulomb_mats,
            orbital_rotations=ham.orbital_rotations, constant=ham.constant,
            expected=linear_operator(ham, norb=norb, nelec=nelec) @ vec)
    # simulate_qdrift_double_factorized (trotter/qdrift.py:23): sampling order, both schedules
    for norb, nelec, rank, z, symmetric, probs, n_steps, n_samples, seed in [
        (4, (2, 2), 3, False, False, "norm", 4, 1, 4101), (4, (2, 2), 3, False, True, "norm", 3, 2, 4102),
        (5, (2, 3), 4, True, False, "uniform", 5, 1, 4103), (4, (1, 2), 2, False, True, "uniform", 2, 1, 4104),
    ]:
        ham = rr.random_double_factorized_hamiltonian(norb, rank=rank, z_representation=z, seed=rng)
        vec = rr.random_state_vector(ref_dim(norb, nelec), seed=rng)
        tag = f"{norb}_{nelec[0]}_{nelec[1]}_r{rank}_{'z' if z else 'num'}_{'sym' if symmetric else 'plain'}_{probs}"
        add(f"qdrift/{tag}", kind="qdrift", norb=norb, nelec=nelec, vec=vec, time=0.3, z=z, symmetric=symmetric,
            probabilities=np.array(list(probs.encode())), n_steps=n_steps, n_samples=n_samples, seed=seed,
            one_body_tensor=ham.one_body_tensor, diag_coulomb_mats=ham.diag_coulomb_mats,
            orbital_rotations=ham.orbital_rotations, constant=ham.constant,
            expected=simulate_qdrift_double_factorized(vec, ham, 0.3, norb=norb, nelec=nelec, n_steps=n_steps,
                                                       symmetric=symmetric, probabilities=probs,
                                                       n_samples=n_samples, seed=seed))

    # generators of the widened operators, explicit seeds (pins ffsim_b200/random.py)
    for fn_name, norb, kw in [
        ("random_ucj_op_spin_unbalanced", 5, dict(n_reps=2, with_final_orbital_rotation=True, seed=5101)),
        ("random_ucj_op_spin_unbalanced", 4, dict(n_reps=1, diag_coulomb_normal=True, diag_coulomb_mean=0.3, seed=5102)),
        ("random_ucj_op_spinless", 5, dict(n_reps=2, with_final_orbital_rotation=True, seed=5103)),
        ("random_ucj_op_spinless", 4, dict(n_reps=3, diag_coulomb_normal=True, seed=5104)),
    ]:
        op = getattr(rr, fn_name)(norb, **kw)
        add(f"random_op/{fn_name}_{kw['seed']}", kind=fn_name, n=norb, seed=kw["seed"], n_reps=kw["n_reps"],
            final=int(kw.get("with_final_orbital_rotation", False)), normal=int(kw.get("diag_coulomb_normal", False)),
            mean=kw.get("diag_coulomb_mean", 0.0), diag_coulomb_mats=op.diag_coulomb_mats,
            orbital_rotations=op.orbital_rotations, final_orbital_rotation=opt(op.final_orbital_rotation))

    # ---- round 2: the rest of SURVEY.md section 8f (appended; earlier random streams unchanged) ----
    from ffsim.gates import basic_gates as bg
    from ffsim.states.spin import Spin
    from ffsim.trotter.qdrift import qdrift_probabilities
    from ffsim.variational.givens import GivensAnsatzOp
    from ffsim.variational.ucj_angles_spin_balanced import UCJAnglesOpSpinBalanced
    from ffsim.variational.ucj_spin_balanced import UCJOpSpinBalanced
    from ffsim.variational.ucj_spin_unbalanced import UCJOpSpinUnbalanced
    from ffsim.variational.ucj_spinless import UCJOpSpinless

    rng2 = np.random.default_rng(20261018)
    spins = {"a": Spin.ALPHA, "b": Spin.BETA, "ab": Spin.ALPHA_AND_BETA}
    # named gates (gates/basic_gates.py:54-629): spinful with every spin choice, and spinless
    for norb, nelec in [(4, (2, 2)), (5, (3, 2)), (5, 2)]:
        vec = rr.random_state_vector(ref_dim(norb, nelec), seed=rng2)
        tag = f"{norb}_{nelec}" if isinstance(nelec, int) else f"{norb}_{nelec[0]}_{nelec[1]}"
        theta, phi = 0.37, -0.81
        for sname, spin in spins.items():
            if isinstance(nelec, int) and sname != "ab":
                continue
            kw = dict(norb=norb, nelec=nelec, spin=spin)
            meta = dict(norb=norb, nelec=nelec, vec=vec, theta=theta, phi=phi, spin=np.array(list(sname.encode())))
            add(f"basic/givens_{tag}_{sname}", kind="basic", gate=np.array(list(b"givens")), orbs=(1, 3), **meta,
                expected=bg.apply_givens_rotation(vec, theta, (1, 3), phi=phi, **kw))
            add(f"basic/tunneling_{tag}_{sname}", kind="basic", gate=np.array(list(b"tunneling")), orbs=(0, 2), **meta,
                expected=bg.apply_tunneling_interaction(vec, theta, (0, 2), **kw))
            add(f"basic/num_{tag}_{sname}", kind="basic", gate=np.array(list(b"num")), orbs=(2, 2), **meta,
                expected=bg.apply_num_interaction(vec, theta, 2, **kw))
            add(f"basic/num_num_{tag}_{sname}", kind="basic", gate=np.array(list(b"num_num")), orbs=(1, 2), **meta,
                expected=bg.apply_num_num_interaction(vec, theta, (1, 2), **kw))
            add(f"basic/hop_{tag}_{sname}", kind="basic", gate=np.array(list(b"hop")), orbs=(3, 1), **meta,
                expected=bg.apply_hop_gate(vec, theta, (3, 1), **kw))
            add(f"basic/fsim_{tag}_{sname}", kind="basic", gate=np.array(list(b"fsim")), orbs=(0, 1), **meta,
                expected=bg.apply_fsim_gate(vec, theta, phi, (0, 1), **kw))
            add(f"basic/fswap_{tag}_{sname}", kind="basic", gate=np.array(list(b"fswap")), orbs=(2, 3), **meta,
                expected=bg.apply_fswap_gate(vec, (2, 3), **kw))
        if not isinstance(nelec, int):
            add(f"basic/on_site_{tag}", kind="basic", gate=np.array(list(b"on_site")), orbs=(1, 1), norb=norb,
                nelec=nelec, vec=vec, theta=theta, phi=phi, spin=np.array(list(b"ab")),
                expected=bg.apply_on_site_interaction(vec, theta, 1, norb=norb, nelec=nelec))
            add(f"basic/num_op_prod_{tag}", kind="basic", gate=np.array(list(b"num_op_prod")), orbs=(0, 2), norb=norb,
                nelec=nelec, vec=vec, theta=theta, phi=phi, spin=np.array(list(b"ab")),
                expected=bg.apply_num_op_prod_interaction(vec, theta, ([0, 2], [1]), norb=norb, nelec=nelec))

    # parameter vectors of the UCJ operators (ucj_spin_balanced.py:144-295 and siblings)
    op = rr.random_ucj_op_spin_balanced(4, n_reps=2, with_final_orbital_rotation=True, seed=6101)
    pairs = ([(0, 1), (1, 2), (2, 3)], [(0, 0), (1, 1), (2, 2), (3, 3)])
    add("params/ucj_balanced_full", kind="params_balanced", norb=4, n_reps=2, final=1, pairs_aa=NONE, pairs_ab=NONE,
        diag_coulomb_mats=op.diag_coulomb_mats, orbital_rotations=op.orbital_rotations,
        final_orbital_rotation=op.final_orbital_rotation, expected=op.to_parameters())
    add("params/ucj_balanced_pairs", kind="params_balanced", norb=4, n_reps=2, final=1, pairs_aa=np.array(pairs[0]),
        pairs_ab=np.array(pairs[1]), diag_coulomb_mats=op.diag_coulomb_mats, orbital_rotations=op.orbital_rotations,
        final_orbital_rotation=op.final_orbital_rotation, expected=op.to_parameters(interaction_pairs=pairs))
    params = rng2.uniform(-1, 1, UCJOpSpinBalanced.n_params(4, 2, interaction_pairs=pairs, with_final_orbital_rotation=True))
    op2 = UCJOpSpinBalanced.from_parameters(params, norb=4, n_reps=2, interaction_pairs=pairs,
                                            with_final_orbital_rotation=True)
    add("params/ucj_balanced_from", kind="params_balanced_from", norb=4, n_reps=2, final=1, pairs_aa=np.array(pairs[0]),
        pairs_ab=np.array(pairs[1]), params=params, diag_coulomb_mats=op2.diag_coulomb_mats,
        orbital_rotations=op2.orbital_rotations, final_orbital_rotation=op2.final_orbital_rotation)
    opu = rr.random_ucj_op_spin_unbalanced(3, n_reps=2, with_final_orbital_rotation=True, seed=6102)
    add("params/ucj_unbalanced_full", kind="params_unbalanced", norb=3, n_reps=2, final=1,
        diag_coulomb_mats=opu.diag_coulomb_mats, orbital_rotations=opu.orbital_rotations,
        final_orbital_rotation=opu.final_orbital_rotation, expected=opu.to_parameters(),
        n_params=UCJOpSpinUnbalanced.n_params(3, 2, with_final_orbital_rotation=True))
    ops = rr.random_ucj_op_spinless(4, n_reps=2, with_final_orbital_rotation=False, seed=6103)
    add("params/ucj_spinless_pairs", kind="params_spinless", norb=4, n_reps=2, final=0, pairs=np.array([(0, 1), (1, 3)]),
        diag_coulomb_mats=ops.diag_coulomb_mats, orbital_rotations=ops.orbital_rotations,
        expected=ops.to_parameters(interaction_pairs=[(0, 1), (1, 3)]),
        n_params=UCJOpSpinless.n_params(4, 2, interaction_pairs=[(0, 1), (1, 3)]))

    # GivensAnsatzOp / UCJAnglesOpSpinBalanced (variational/givens.py, ucj_angles_spin_balanced.py)
    u = rr.random_unitary(5, seed=6104)
    g = GivensAnsatzOp.from_orbital_rotation(u)
    add("angles/givens_from_rotation_5", kind="givens_from_rotation", norb=5, mat=u,
        pairs=np.array(g.interaction_pairs), thetas=g.thetas, phis=g.phis, phase_angles=g.phase_angles,
        rotation=g.to_orbital_rotation())
    for norb, nelec, n_reps, final in [(4, (2, 2), 2, True), (5, (2, 3), 1, False)]:
        nn_pairs = ([(p, p + 1) for p in range(norb - 1)], [(p, p) for p in range(norb)])
        g_pairs = [(p, p + 1) for p in range(0, norb - 1, 2)] + [(p, p + 1) for p in range(1, norb - 1, 2)]
        n_par = UCJAnglesOpSpinBalanced.n_params(norb, n_reps, nn_pairs, g_pairs, with_final_givens_ansatz_op=final)
        params = rng2.uniform(-np.pi, np.pi, n_par)
        aop = UCJAnglesOpSpinBalanced.from_parameters(params, norb=norb, n_reps=n_reps, num_num_interaction_pairs=nn_pairs,
                                                      givens_interaction_pairs=g_pairs, with_final_givens_ansatz_op=final)
        vec = rr.random_state_vector(ref_dim(norb, nelec), seed=rng2)
        add(f"angles/ucj_angles_{norb}_{nelec[0]}_{nelec[1]}_L{n_reps}", kind="ucj_angles", norb=norb, nelec=nelec,
            n_reps=n_reps, final=int(final), params=params, pairs_aa=np.array(nn_pairs[0]), pairs_ab=np.array(nn_pairs[1]),
            givens_pairs=np.array(g_pairs), vec=vec, roundtrip=aop.to_parameters(),
            expected=apply_unitary(vec, aop, norb=norb, nelec=nelec))
    ucj = rr.random_ucj_op_spin_balanced(4, n_reps=2, with_final_orbital_rotation=True, seed=6105)
    aop = UCJAnglesOpSpinBalanced.from_ucj_op(ucj)
    vec = rr.random_state_vector(ref_dim(4, (2, 2)), seed=rng2)
    add("angles/from_ucj_op_4", kind="ucj_angles_from_ucj", norb=4, nelec=(2, 2), vec=vec,
        diag_coulomb_mats=ucj.diag_coulomb_mats, orbital_rotations=ucj.orbital_rotations,
        final_orbital_rotation=ucj.final_orbital_rotation, params=aop.to_parameters(),
        expected=apply_unitary(vec, aop, norb=4, nelec=(2, 2)))

    # qDRIFT sampling probabilities incl. the state-dependent ones (trotter/qdrift.py:247-348, states/wick.py)
    for norb, nelec, rank, z in [(4, (2, 2), 3, False), (4, (1, 2), 2, True)]:
        ham = rr.random_double_factorized_hamiltonian(norb, rank=rank, z_representation=z, seed=rng2)
        # one-rdm of a Slater determinant (not spin-summed, spin-orbital basis): occupied orbitals of a random basis
        import scipy.linalg
        ua, ub = rr.random_unitary(norb, seed=rng2), rr.random_unitary(norb, seed=rng2)
        rdm = scipy.linalg.block_diag((ua[:, : nelec[0]] @ ua[:, : nelec[0]].T.conj()).T,
                                      (ub[:, : nelec[1]] @ ub[:, : nelec[1]].T.conj()).T)
        tag = f"{norb}_{nelec[0]}_{nelec[1]}_r{rank}_{'z' if z else 'num'}"
        for method in ("norm", "uniform", "optimal", "optimal-incoherent"):
            add(f"qdrift_probs/{method}_{tag}", kind="qdrift_probs", norb=norb, nelec=nelec, z=z,
                method=np.array(list(method.encode())), one_rdm=rdm, one_body_tensor=ham.one_body_tensor,
                diag_coulomb_mats=ham.diag_coulomb_mats, orbital_rotations=ham.orbital_rotations, constant=ham.constant,
                expected=qdrift_probabilities(ham, sampling_method=method, nelec=nelec, one_rdm=rdm))

    flat = {}
    for name, arrays in CASES.items():
        for k, v in arrays.items():
            flat[f"{name}::{k}"] = v
    out = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(out, **flat)
    print(f"wrote {len(CASES)} cases, {os.path.getsize(out) / 1024:.1f} KiB -> {out}")


if __name__ == "__main__":
    main()
