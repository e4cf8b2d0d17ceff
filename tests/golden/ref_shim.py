"""Import shim that runs the REFERENCE's own Python code in this container.

ffsim cannot be imported here as a package: its Rust extension (`ffsim._lib`) is not built
(no cargo) and pyscf / jax / qiskit / opt_einsum / orjson are not installed.  This shim
loads the reference's Python modules from /root/reference/python/ffsim *file by file*,
unmodified, and supplies only what is missing:

* `ffsim._lib`   -> the reference's own pure-Python twins of the Rust kernels
                    (python/ffsim/_slow/**; the reference's tests assert twin == Rust,
                    e.g. tests/python/_slow/gates/orbital_rotation_test.py), plus
                    `givens_decomposition` from oracle/givens.py (restatement of
                    src/linalg/givens.rs:20-149 -- there is no Python twin of it) and a
                    three-line `apply_phase_shift_in_place` (src/gates/phase_shift.rs:18-30);
* `pyscf.fci.cistring` -> oracle/cistring.py (restatement);
* jax, qiskit, opt_einsum, orjson, other pyscf modules -> inert stubs (never called on the
  hot path).

Used only by tests/golden/make_golden.py, here, to write fixtures.  Nothing on the GPU box
imports this file.
"""

from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

REF_PKG = "/root/reference/python/ffsim"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _Dummy:
    """Inert placeholder: usable as a decorator, a base class argument or an attribute bag."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]  # decorator use: @jax.jit
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __mro_entries__(self, bases):
        return (object,)

    def __getitem__(self, item):
        return _Dummy()

    def __or__(self, other):
        return _Dummy()

    __ror__ = __or__


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


STUB_ROOTS = ("jax", "qiskit", "opt_einsum", "orjson", "pyscf")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in STUB_ROOTS and fullname != "pyscf.fci.cistring":
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def install():
    """Make `import ffsim.<sub>` resolve to the reference's files, with the stand-ins above."""
    if "ffsim" in sys.modules and getattr(sys.modules["ffsim"], "__ref_shim__", False):
        return sys.modules["ffsim"]
    assert os.path.isdir(REF_PKG), "the reference tree is only mounted in the build container"
    sys.meta_path.insert(0, _StubFinder())

    from oracle import cistring as ocis
    from oracle import givens as ogivens

    # opt_einsum.contract is numpy.einsum with a different contraction-order optimiser
    import numpy as _np
    import opt_einsum as _oe  # (stub module)

    _oe.contract = lambda subscripts, *operands, **kw: _np.einsum(subscripts, *operands, optimize=True)
    # scipy's array-API helpers ask `issubclass(type(x), sys.modules["jax"].Array)` once "jax" is in
    # sys.modules: give the stub a real (empty) class there
    import jax as _jax  # (stub module)

    _jax.Array = type("Array", (), {})

    # pyscf.fci.cistring
    import pyscf.fci  # noqa: F401  (stub)

    cis = types.ModuleType("pyscf.fci.cistring")
    cis.make_strings = ocis.make_strings
    cis.gen_strings4orblist = ocis.make_strings
    cis.gen_occslst = lambda orbs, n: ocis.gen_occslst(orbs, n).astype("int32")
    cis.str2addr = lambda norb, nelec, s: int(ocis.strs2addr(norb, nelec, [s])[0])
    cis.strs2addr = ocis.strs2addr
    cis.num_strings = lambda n, m: __import__("math").comb(n, m)
    sys.modules["pyscf.fci.cistring"] = cis
    sys.modules["pyscf.fci"].cistring = cis

    # the package shell: no __init__ executed (it pulls in qiskit, the Rust operators, ...)
    pkg = types.ModuleType("ffsim")
    pkg.__path__ = [REF_PKG]
    pkg.__ref_shim__ = True
    sys.modules["ffsim"] = pkg

    # ffsim._lib from the reference's _slow twins; basic_gates (needed by one twin) imports
    # ffsim._lib itself, so install the module first and fill it in afterwards
    lib = _StubModule("ffsim._lib")  # unknown names (operator kernels etc.) resolve to inert dummies
    sys.modules["ffsim._lib"] = lib
    pkg._lib = lib
    lib.FermionOperator = type("FermionOperator", (), {})  # Rust class; only used in isinstance checks here
    lib.givens_decomposition = lambda mat, tol=1e-12: ogivens.givens_decomposition(mat, tol)

    def apply_phase_shift_in_place(vec, phase, indices):
        # src/gates/phase_shift.rs:18-30 (three lines; the reference ships no Python twin of it)
        for i in indices:
            vec[int(i)] *= phase

    lib.apply_phase_shift_in_place = apply_phase_shift_in_place

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_PKG, rel))
        mod = importlib.util.module_from_spec(spec)
        mod.__dict__.setdefault("__package__", name.rpartition(".")[0])
        spec.loader.exec_module(mod)
        return mod

    slow_or = load("_ref_slow_orbital_rotation", "_slow/gates/orbital_rotation.py")
    slow_no = load("_ref_slow_num_op_sum", "_slow/gates/num_op_sum.py")
    slow_cd = load("_ref_slow_contract_dc", "_slow/contract/diag_coulomb.py")
    slow_cn = load("_ref_slow_contract_nos", "_slow/contract/num_op_sum.py")
    lib.apply_givens_rotation_in_place = slow_or.apply_givens_rotation_in_place_slow
    lib.apply_num_op_sum_evolution_in_place = slow_no.apply_num_op_sum_evolution_in_place_slow
    lib.contract_diag_coulomb_into_buffer_num_rep = slow_cd.contract_diag_coulomb_into_buffer_num_rep_slow
    lib.contract_diag_coulomb_into_buffer_z_rep = slow_cd.contract_diag_coulomb_into_buffer_z_rep_slow
    lib.contract_num_op_sum_spin_into_buffer = slow_cn.contract_num_op_sum_spin_into_buffer_slow
    # this twin imports ffsim.gates.basic_gates -> ffsim.gates -> `from ffsim._lib import
    # apply_diag_coulomb_evolution_in_place_*`: bind those names late
    holder = {}
    lib.apply_diag_coulomb_evolution_in_place_num_rep = lambda *a, **k: holder[
        "dc"].apply_diag_coulomb_evolution_in_place_num_rep_slow(*a, **k)
    lib.apply_diag_coulomb_evolution_in_place_z_rep = lambda *a, **k: holder[
        "dc"].apply_diag_coulomb_evolution_in_place_z_rep_slow(*a, **k)
    holder["dc"] = load("_ref_slow_diag_coulomb", "_slow/gates/diag_coulomb.py")
    return pkg
