"""Gates on the hot path, and the named two-orbital gates built on them."""

from ffsim_b200.gates.basic_gates import (
    apply_fsim_gate,
    apply_fswap_gate,
    apply_givens_rotation,
    apply_hop_gate,
    apply_num_interaction,
    apply_num_num_interaction,
    apply_num_op_prod_interaction,
    apply_on_site_interaction,
    apply_tunneling_interaction,
)
from ffsim_b200.gates.diag_coulomb import apply_diag_coulomb_evolution
from ffsim_b200.gates.num_op_sum import apply_num_op_sum_evolution
from ffsim_b200.gates.orbital_rotation import apply_orbital_rotation

__all__ = [
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
]
