"""Diagonal contractions on the hot path."""

from ffsim_b200.contract.diag_coulomb import contract_diag_coulomb, diag_coulomb_linop
from ffsim_b200.contract.linop import DeviceLinearOperator
from ffsim_b200.contract.num_op_sum import contract_num_op_sum, num_op_sum_linop

__all__ = [
    "DeviceLinearOperator",
    "contract_diag_coulomb",
    "contract_num_op_sum",
    "diag_coulomb_linop",
    "num_op_sum_linop",
]
