from ffsim_b200.trotter.diagonal_coulomb_split_op import simulate_trotter_diag_coulomb_split_op
from ffsim_b200.trotter.double_factorized import simulate_trotter_double_factorized
from ffsim_b200.trotter.qdrift import qdrift_probabilities, simulate_qdrift_double_factorized

__all__ = [
    "qdrift_probabilities",
    "simulate_qdrift_double_factorized",
    "simulate_trotter_diag_coulomb_split_op",
    "simulate_trotter_double_factorized",
]
