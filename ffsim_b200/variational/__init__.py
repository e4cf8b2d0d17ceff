from ffsim_b200.variational.ucj_spin_balanced import UCJOpSpinBalanced
from ffsim_b200.variational.ucj_spin_unbalanced import UCJOpSpinUnbalanced
from ffsim_b200.variational.ucj_spinless import UCJOpSpinless

__all__ = ["UCJOpSpinBalanced", "UCJOpSpinUnbalanced", "UCJOpSpinless"]
