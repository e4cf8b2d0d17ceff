"""ctypes binding of the C ABI in ``include/ffsim_b200.h`` (libffsim_b200.so).

This is the counterpart of the reference's ``ffsim._lib`` PyO3 module
(src/lib.rs:20-69).  There is no fallback: if the shared library has not been
built (``python -c 'import __graft_entry__ as g; g.build()'``) importing this
module raises ImportError.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFSIM_B200_LIB") or os.path.join(_HERE, "lib", "libffsim_b200.so")  # env: developer override

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA library first "
        "(python -c 'import __graft_entry__ as g; g.build()' or make -C ffsim_b200/csrc). "
        "ffsim_b200 has no CPU fallback."
    )

lib = ctypes.CDLL(LIB_PATH)

FFB_OK = 0
FFB_EINVAL = -1
FFB_ECUDA = -2
FFB_ENOMEM = -3
FFB_EINTERNAL = -4


class C128(ctypes.Structure):
    _fields_ = [("re", c_double), ("im", c_double)]


class GivensRotation(ctypes.Structure):
    _fields_ = [("c", c_double), ("s", C128), ("i", c_int32), ("j", c_int32)]


GIVENS_DTYPE = np.dtype(
    {"names": ["c", "s", "i", "j"], "formats": [np.float64, np.complex128, np.int32, np.int32],
     "offsets": [0, 8, 24, 28], "itemsize": 32}
)
assert ctypes.sizeof(GivensRotation) == GIVENS_DTYPE.itemsize

_P = c_void_p


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


EXPORTS = {
    "ffb_version": (c_int,),
    "ffb_last_error": (c_char_p,),
    "ffb_device_count": (c_int,),
    "ffb_tables_create": (c_int, c_int, c_int, POINTER(_P)),
    "ffb_tables_destroy": (None, _P),
    "ffb_tables_dim": (c_int64, _P),
    "ffb_tables_norb": (c_int, _P),
    "ffb_tables_nocc": (c_int, _P),
    "ffb_tables_strings": (c_int, _P, _P),
    "ffb_tables_occupations": (c_int, _P, _P),
    "ffb_tables_strs2addr": (c_int, _P, _P, c_int64, _P),
    "ffb_tables_n_pairs": (c_int64, _P),
    "ffb_tables_zero_one_subspace": (c_int, _P, c_int, c_int, _P, POINTER(c_int64)),
    "ffb_tables_n_one": (c_int64, _P),
    "ffb_tables_one_subspace": (c_int, _P, c_int, _P, POINTER(c_int64)),
    "ffb_givens_decomposition": (c_int, _P, c_int, c_double, _P, POINTER(c_int), _P),
    "ffb_apply_givens_rotation_in_place": (c_int, _P, c_int64, c_int64, c_int64, c_double, C128, _P, _P, c_int64, _P),
    "ffb_apply_phase_shift_in_place": (c_int, _P, c_int64, c_int64, c_int64, C128, _P, c_int64, _P),
    "ffb_plan_orbital_rotation": (c_int, _P, _P, _P, c_int, _P, _P, c_int, _P, POINTER(_P)),
    "ffb_plan_destroy": (None, _P),
    "ffb_plan_update_coefficients": (c_int, _P, _P, c_int, _P, _P, c_int, _P),
    "ffb_set_device": (c_int, c_int),
    "ffb_plan_workspace_bytes": (c_int64, _P, c_int64),
    "ffb_plan_describe": (c_int, _P, c_char_p, c_size_t),
    "ffb_plan_n_state_passes": (c_int, _P),
    "ffb_apply_orbital_rotation": (c_int, _P, _P, _P, _P),
    "ffb_apply_orbital_rotation_rows": (c_int, _P, c_int, _P, c_int64, c_int64, _P),
    "ffb_apply_orbital_rotation_strided": (c_int, _P, c_int, _P, c_int64, c_int64, c_int64, _P),
    "ffb_plan_beta_in_place": (c_int, _P),
    "ffb_apply_orbital_rotation_beta_block": (c_int, _P, _P, c_int64, c_int64, _P, _P),
    "ffb_apply_diag_coulomb_evolution": (c_int, _P, _P, _P, _P, _P, c_int, _P, c_int64, c_int64, _P),
    "ffb_apply_num_op_sum_evolution": (c_int, _P, _P, _P, _P, _P, c_int64, c_int64, _P),
    "ffb_contract_diag_coulomb": (c_int, _P, _P, _P, _P, _P, c_int, _P, _P, c_int, c_int64, c_int64, _P),
    "ffb_contract_num_op_sum": (c_int, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int64, _P),
    "ffb_apply_diag_coulomb_evolution_block": (c_int, _P, _P, _P, _P, _P, c_int, _P, c_int64, c_int64, c_int64, c_int64, c_int64, _P),
    "ffb_apply_num_op_sum_evolution_block": (c_int, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, _P),
    "ffb_contract_diag_coulomb_block": (c_int, _P, _P, _P, _P, _P, c_int, _P, _P, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, _P),
    "ffb_contract_num_op_sum_block": (c_int, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, _P),
    "ffb_apply_num_op_prod_phase": (c_int, _P, _P, ctypes.c_uint32, ctypes.c_uint32, C128, _P, c_int64, c_int64, c_int64,
                                    c_int64, c_int64, _P),
    "ffb_transpose": (c_int, _P, _P, c_int64, c_int64, c_int64, c_int64, _P),
    "ffb_exchange_blocks": (c_int, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P),
    "ffb_copy_blocks": (c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P),
    "ffb_vdot": (c_int, _P, _P, c_int64, _P, _P),
    "ffb_axpby": (c_int, C128, _P, C128, _P, c_int64, _P),
    "ffb_profile_begin": (c_int,),
    "ffb_profile_end": (c_int, c_char_p, c_size_t),
    "ffb_memcpy2d_async": (c_int, _P, c_size_t, _P, c_size_t, c_size_t, c_size_t, c_int, _P),
    "ffb_measure_fp64_peak": (c_int, POINTER(c_double)),
    "ffb_set_option": (c_int, c_char_p, c_int64),
    "ffb_get_option": (c_int64, c_char_p),
}

for _name, (_res, *_args) in EXPORTS.items():
    _sig(_name, _res, *_args)


class FfbError(RuntimeError):
    pass


def last_error() -> str:
    msg = lib.ffb_last_error()
    return msg.decode() if msg else ""


def check(rc: int) -> None:
    """Map a C status to the exception the reference raises at this boundary."""
    if rc == FFB_OK:
        return
    msg = last_error()
    if rc == FFB_EINVAL:
        raise ValueError(msg)
    if rc == FFB_ENOMEM:
        raise MemoryError(msg)
    raise FfbError(f"ffsim_b200 error {rc}: {msg}")


def ptr(arr: np.ndarray | None):
    return None if arr is None else arr.ctypes.data_as(c_void_p)


def c128(z: complex) -> C128:
    z = complex(z)
    return C128(z.real, z.imag)


def profile_begin() -> None:
    check(lib.ffb_profile_begin())


def profile_end() -> dict:
    import json

    buf = ctypes.create_string_buffer(4096)
    check(lib.ffb_profile_end(buf, len(buf)))
    return json.loads(buf.value.decode())


def measure_fp64_peak() -> float:
    """Dense DFMA throughput of the current device (TFLOP/s), measured now."""
    out = c_double(0.0)
    check(lib.ffb_measure_fp64_peak(byref(out)))
    return float(out.value)


def set_option(key: str, value: int) -> None:
    check(lib.ffb_set_option(key.encode(), int(value)))


def get_option(key: str) -> int:
    return int(lib.ffb_get_option(key.encode()))
