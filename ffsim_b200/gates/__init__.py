"""Gates on the hot path."""

from ffsim_b200.gates.diag_coulomb import apply_diag_coulomb_evolution
from ffsim_b200.gates.num_op_sum import apply_num_op_sum_evolution
from ffsim_b200.gates.orbital_rotation import apply_orbital_rotation

__all__ = ["apply_diag_coulomb_evolution", "apply_num_op_sum_evolution", "apply_orbital_rotation"]
