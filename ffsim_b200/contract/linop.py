"""A SciPy LinearOperator whose matvec also accepts CUDA tensors.

``scipy.sparse.linalg.LinearOperator`` coerces its argument with ``np.asanyarray``,
which a CUDA tensor cannot pass through, so ``@`` / ``matvec`` / ``dot`` are
routed around it for tensors.  NumPy input takes the stock SciPy path (and is
uploaded inside ``_matvec``), so ``eigsh``, ``expm_multiply`` etc. work unchanged.
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import scipy.sparse.linalg
import torch

from ffsim_b200 import _device


class DeviceLinearOperator(scipy.sparse.linalg.LinearOperator):
    def __init__(self, dim: int, device_matvec: Callable[[torch.Tensor], torch.Tensor]):
        super().__init__(dtype=np.dtype(complex), shape=(dim, dim))
        self._device_matvec = device_matvec  # maps a 1-D CUDA tensor to a NEW 1-D CUDA tensor

    # --- tensors
    def matvec_device(self, vec):
        if vec.numel() != self.shape[1]:
            raise ValueError(f"dimension mismatch: {vec.numel()} vs {self.shape[1]}")
        t, kind = _device.to_device(vec, copy=False)
        return _device.from_device(self._device_matvec(t), kind)

    def matvec(self, x):
        if isinstance(x, torch.Tensor) or _device.is_sharded(x):
            return self.matvec_device(x)
        return super().matvec(x)

    rmatvec_device = matvec_device  # Hermitian operators only (rmatvec=matvec in the reference)

    def dot(self, x):
        if isinstance(x, torch.Tensor) or _device.is_sharded(x):
            return self.matvec_device(x)
        return super().dot(x)

    def __matmul__(self, x):
        if isinstance(x, torch.Tensor) or _device.is_sharded(x):
            return self.matvec_device(x)
        return super().__matmul__(x)

    def __call__(self, x):
        return self @ x

    # --- NumPy (SciPy's own entry points end up here)
    def _matvec(self, x):
        arr = np.asarray(x).reshape(-1)
        t, _ = _device.to_device(arr, copy=False)
        return self._device_matvec(t).cpu().numpy()

    def _rmatvec(self, x):
        return self._matvec(x)

    def _adjoint(self):
        return self
