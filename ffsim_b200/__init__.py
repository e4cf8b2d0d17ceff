"""ffsim_b200: the determinant-space statevector hot path of ffsim on B200.

Drop-in for the following names of ``ffsim`` (same signatures and semantics; see
SURVEY.md section 8b).  Vectors may be NumPy arrays (uploaded, result downloaded)
or CUDA ``torch.complex128`` tensors (stay on the device).  All arithmetic on the
state runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/ffsim_b200.h``; there is no CPU fallback.
"""

from ffsim_b200 import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from ffsim_b200 import contract, linalg, random
from ffsim_b200.contract import contract_diag_coulomb, contract_num_op_sum, diag_coulomb_linop, num_op_sum_linop
from ffsim_b200.gates import apply_diag_coulomb_evolution, apply_num_op_sum_evolution, apply_orbital_rotation
from ffsim_b200.hamiltonians import DiagonalCoulombHamiltonian, DoubleFactorizedHamiltonian
from ffsim_b200.init_cache import init_cache
from ffsim_b200.protocols import apply_unitary, linear_operator
from ffsim_b200.states import dim, dims, hartree_fock_state
from ffsim_b200.trotter import (
    qdrift_probabilities,
    simulate_qdrift_double_factorized,
    simulate_trotter_diag_coulomb_split_op,
    simulate_trotter_double_factorized,
)
from ffsim_b200.variational import UCJOpSpinBalanced, UCJOpSpinless, UCJOpSpinUnbalanced

__version__ = "0.1.0"

__all__ = [
    "DiagonalCoulombHamiltonian",
    "DoubleFactorizedHamiltonian",
    "UCJOpSpinBalanced",
    "UCJOpSpinUnbalanced",
    "UCJOpSpinless",
    "apply_diag_coulomb_evolution",
    "apply_num_op_sum_evolution",
    "apply_orbital_rotation",
    "apply_unitary",
    "contract",
    "contract_diag_coulomb",
    "contract_num_op_sum",
    "diag_coulomb_linop",
    "dim",
    "dims",
    "hartree_fock_state",
    "init_cache",
    "linalg",
    "linear_operator",
    "num_op_sum_linop",
    "qdrift_probabilities",
    "random",
    "simulate_qdrift_double_factorized",
    "simulate_trotter_diag_coulomb_split_op",
    "simulate_trotter_double_factorized",
]
