from ffsim_b200.variational.ucj_spin_balanced import UCJOpSpinBalanced

__all__ = ["UCJOpSpinBalanced"]
