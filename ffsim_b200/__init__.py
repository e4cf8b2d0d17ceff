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
from ffsim_b200.gates import (
    apply_diag_coulomb_evolution,
    apply_fsim_gate,
    apply_fswap_gate,
    apply_givens_rotation,
    apply_hop_gate,
    apply_num_interaction,
    apply_num_num_interaction,
    apply_num_op_prod_interaction,
    apply_num_op_sum_evolution,
    apply_on_site_interaction,
    apply_orbital_rotation,
    apply_tunneling_interaction,
)
from ffsim_b200.hamiltonians import DiagonalCoulombHamiltonian, DoubleFactorizedHamiltonian
from ffsim_b200.init_cache import init_cache
from ffsim_b200.pipeline import (
    HostEvolution,
    evolve_host,
    evolve_host_async,
    evolve_host_many,
    evolve_host_rows_async,
    pinned_empty,
    release_device_buffers,
)
from ffsim_b200.protocols import apply_unitary, linear_operator
from ffsim_b200.states import Spin, dim, dims, hartree_fock_state
from ffsim_b200.trotter import (
    qdrift_probabilities,
    simulate_qdrift_double_factorized,
    simulate_trotter_diag_coulomb_split_op,
    simulate_trotter_double_factorized,
)
from ffsim_b200.variational import (
    GivensAnsatzOp,
    NumNumAnsatzOpSpinBalanced,
    UCJAnglesOpSpinBalanced,
    UCJOpSpinBalanced,
    UCJOpSpinless,
    UCJOpSpinUnbalanced,
)

__version__ = "0.2.0"


def to_device(vec):
    """Upload a host vector (NumPy array or CPU tensor) once: a 1-D ``complex128`` CUDA tensor that the
    functions of this package update in place with ``copy=False``.  Pinned host memory goes straight to
    the DMA engine; pageable memory is staged through a pooled pinned buffer."""
    from ffsim_b200 import _device

    return _device.to_device(vec, copy=True)[0]


def to_host(t):
    """Download a CUDA tensor produced by this package: a NumPy array over (pooled) pinned memory."""
    from ffsim_b200 import _device

    if _device.is_sharded(t):
        raise TypeError("to_host: pass the shard (ShardedVector.local), not the distributed vector")
    return _device._download(t.reshape(-1))

__all__ = [
    "DiagonalCoulombHamiltonian",
    "DoubleFactorizedHamiltonian",
    "GivensAnsatzOp",
    "NumNumAnsatzOpSpinBalanced",
    "Spin",
    "UCJAnglesOpSpinBalanced",
    "UCJOpSpinBalanced",
    "UCJOpSpinUnbalanced",
    "UCJOpSpinless",
    "apply_diag_coulomb_evolution",
    "apply_fsim_gate",
    "apply_fswap_gate",
    "apply_givens_rotation",
    "apply_hop_gate",
    "apply_num_interaction",
    "apply_num_num_interaction",
    "apply_num_op_prod_interaction",
    "apply_num_op_sum_evolution",
    "apply_on_site_interaction",
    "apply_orbital_rotation",
    "apply_tunneling_interaction",
    "apply_unitary",
    "contract",
    "contract_diag_coulomb",
    "contract_num_op_sum",
    "diag_coulomb_linop",
    "HostEvolution",
    "evolve_host",
    "evolve_host_async",
    "evolve_host_many",
    "evolve_host_rows_async",
    "release_device_buffers",
    "dim",
    "dims",
    "hartree_fock_state",
    "init_cache",
    "linalg",
    "linear_operator",
    "num_op_sum_linop",
    "pinned_empty",
    "qdrift_probabilities",
    "random",
    "simulate_qdrift_double_factorized",
    "simulate_trotter_diag_coulomb_split_op",
    "simulate_trotter_double_factorized",
    "to_device",
    "to_host",
]
