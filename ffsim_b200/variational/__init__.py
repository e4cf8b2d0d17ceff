from ffsim_b200.variational.givens import GivensAnsatzOp
from ffsim_b200.variational.num_num import NumNumAnsatzOpSpinBalanced
from ffsim_b200.variational.ucj_angles_spin_balanced import UCJAnglesOpSpinBalanced
from ffsim_b200.variational.ucj_spin_balanced import UCJOpSpinBalanced
from ffsim_b200.variational.ucj_spin_unbalanced import UCJOpSpinUnbalanced
from ffsim_b200.variational.ucj_spinless import UCJOpSpinless

__all__ = ["GivensAnsatzOp", "NumNumAnsatzOpSpinBalanced", "UCJAnglesOpSpinBalanced", "UCJOpSpinBalanced",
           "UCJOpSpinUnbalanced", "UCJOpSpinless"]
