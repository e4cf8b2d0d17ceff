"""CPU restatement of the named two-orbital gates and the angle-parameterised ansatz operators.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows python/ffsim/gates/basic_gates.py:27-629,
python/ffsim/variational/givens.py:203-222, num_num.py:158-175 and ucj_angles_spin_balanced.py:183-222
over the oracle's own ``apply_orbital_rotation`` / ``apply_num_op_sum_evolution`` /
``apply_diag_coulomb_evolution``; pinned against reference-generated vectors in
tests/test_reference_golden.py.
"""

from __future__ import annotations

import cmath
import math

import numpy as np

from oracle import gates
from oracle.cistring import make_strings


def _pair(obj, spin: str):
    """spin: 'a', 'b' or 'ab' (ffsim.Spin.ALPHA / BETA / ALPHA_AND_BETA)."""
    return (obj if "a" in spin else None, obj if "b" in spin else None)


def _two_orbital(norb, orbs, block):
    mat = np.eye(norb, dtype=complex)
    mat[np.ix_(orbs, orbs)] = block
    return mat


def apply_givens_rotation(vec, theta, orbs, norb, nelec, spin="ab", phi=0.0):
    c, s = math.cos(theta), cmath.exp(1j * phi) * math.sin(theta)
    mat = _two_orbital(norb, orbs, [[c, s], [-s.conjugate(), c]])
    return gates.apply_orbital_rotation(vec, mat if isinstance(nelec, int) else _pair(mat, spin), norb, nelec)


def apply_num_interaction(vec, theta, orb, norb, nelec, spin="ab"):
    coeffs = np.zeros(norb)
    coeffs[orb] = 1
    return gates.apply_num_op_sum_evolution(vec, coeffs if isinstance(nelec, int) else _pair(coeffs, spin), -theta,
                                            norb, nelec)


def apply_tunneling_interaction(vec, theta, orbs, norb, nelec, spin="ab"):
    vec = apply_num_interaction(vec, -math.pi / 2, orbs[0], norb, nelec, spin)
    vec = apply_givens_rotation(vec, theta, orbs, norb, nelec, spin)
    return apply_num_interaction(vec, math.pi / 2, orbs[0], norb, nelec, spin)


def apply_num_op_prod_interaction(vec, theta, target_orbs, norb, nelec):
    """basic_gates.py:27-51,375-424: phase on amplitudes whose strings contain all target orbitals."""
    alpha_orbs, beta_orbs = target_orbs
    sa, sb = make_strings(range(norb), nelec[0]), make_strings(range(norb), nelec[1])
    ma = sum(1 << int(p) for p in alpha_orbs)
    mb = sum(1 << int(p) for p in beta_orbs)
    rows = np.array([(int(s) & ma) == ma for s in sa])
    cols = np.array([(int(s) & mb) == mb for s in sb])
    out = np.array(vec, dtype=complex).reshape(len(sa), len(sb)).copy()
    out[np.ix_(rows, cols)] *= cmath.exp(1j * theta)
    return out.reshape(-1)


def apply_num_num_interaction(vec, theta, orbs, norb, nelec, spin="ab"):
    if isinstance(nelec, int):
        return apply_num_op_prod_interaction(vec, theta, (orbs, []), norb, (nelec, 0))
    if "a" in spin:
        vec = apply_num_op_prod_interaction(vec, theta, (orbs, []), norb, nelec)
    if "b" in spin:
        vec = apply_num_op_prod_interaction(vec, theta, ([], orbs), norb, nelec)
    return np.array(vec, dtype=complex)


def apply_on_site_interaction(vec, theta, orb, norb, nelec):
    return apply_num_op_prod_interaction(vec, theta, ([orb], [orb]), norb, nelec)


def apply_hop_gate(vec, theta, orbs, norb, nelec, spin="ab"):
    return apply_num_num_interaction(apply_givens_rotation(vec, theta, orbs, norb, nelec, spin), math.pi, orbs, norb,
                                     nelec, spin)


def apply_fsim_gate(vec, theta, phi, orbs, norb, nelec, spin="ab"):
    return apply_num_num_interaction(apply_tunneling_interaction(vec, -theta, orbs, norb, nelec, spin), -phi, orbs,
                                     norb, nelec, spin)


def apply_fswap_gate(vec, orbs, norb, nelec, spin="ab"):
    mat = _two_orbital(norb, orbs, [[0, 1], [1, 0]])
    return gates.apply_orbital_rotation(vec, mat if isinstance(nelec, int) else _pair(mat, spin), norb, nelec)


# ---------------------------------------------------------------- angle-parameterised ansatz operators
def givens_ansatz_rotation(norb, pairs, thetas, phis, phase_angles):
    """GivensAnsatzOp.to_orbital_rotation (variational/givens.py:203-222)."""
    u = np.diag(np.exp(1j * np.asarray(phase_angles, dtype=float))).astype(complex)
    for (i, j), theta, phi in zip(list(pairs)[::-1], list(thetas)[::-1], list(phis)[::-1]):
        c, s = math.cos(theta), cmath.rect(math.sin(theta), -phi)
        x, y = u[:, j].copy(), u[:, i].copy()
        u[:, j], u[:, i] = c * x + s * y, c * y - np.conj(s) * x
    return u


def ucj_angles_apply(vec, norb, nelec, n_reps, params, pairs_aa, pairs_ab, givens_pairs, with_final):
    """UCJAnglesOpSpinBalanced.from_parameters + _apply_unitary_ (ucj_angles_spin_balanced.py:69-135,183-222)."""
    ng, naa, nab = len(givens_pairs), len(pairs_aa), len(pairs_ab)
    pos, basis = 0, np.eye(norb)
    out = np.array(vec, dtype=complex)
    for _ in range(n_reps):
        thetas, phis, phase = params[pos : pos + ng], params[pos + ng : pos + 2 * ng], params[pos + 2 * ng : pos + 2 * ng + norb]
        pos += 2 * ng + norb
        rot = givens_ansatz_rotation(norb, givens_pairs, thetas, phis, phase)
        mats = np.zeros((2, norb, norb))
        for k, pairs in enumerate((pairs_aa, pairs_ab)):
            for (i, j), theta in zip(pairs, params[pos : pos + len(pairs)]):
                mats[k, i, j] = mats[k, j, i] = theta
            pos += len(pairs)
        out = gates.apply_orbital_rotation(out, rot.T.conj() @ basis, norb, nelec)
        out = gates.apply_diag_coulomb_evolution(out, (mats[0], mats[1], mats[0]), -1.0, norb, nelec)
        basis = rot
    if with_final:
        brick = [(j, j + 1) for layer in range(norb) for j in range(layer % 2, norb - 1, 2)]
        nb = len(brick)
        rest = params[pos:]
        basis = givens_ansatz_rotation(norb, brick, rest[:nb], rest[nb : 2 * nb], rest[2 * nb :]) @ basis
    return gates.apply_orbital_rotation(out, basis, norb, nelec)
