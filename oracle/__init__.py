"""CPU oracle for the ffsim determinant-space statevector hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ffsim_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline — never as the product path.

What it is: a numpy restatement of the reference algorithm (ffsim 0.0.85.dev,
Python drivers + Rust kernels) for the path SURVEY.md section 8 names.  Every
function cites the reference file:line it follows (paths relative to
``/root/reference``).  The reference itself cannot be imported or built in this
environment (no pyscf / jax / qiskit / cargo), so the oracle is pinned against
the golden values the reference tree holds for this path instead:

* docs/explanations/state-vectors-and-gates.ipynb cells 9, 11, 13
* docs/explanations/diag-coulomb-hamiltonian.ipynb cells 5, 7
* tests/python/states/bitstring_test.py:24-97 (string tables)
* python/ffsim/states/bitstring.py docstring examples
* an independent closed form (compound matrices / Slater minors, ``compound.py``)

``tests/test_oracle_golden.py`` checks all of them: parity is PINNED for the
orbital rotation, diagonal Coulomb evolution / contraction, number-operator-sum
evolution / contraction, the DiagonalCoulombHamiltonian linear operator and the
split-operator Trotter driver.

Third-party arithmetic that is not in /root/reference: ``pyscf.fci.cistring``
(pinned pyscf 2.14.0 in uv.lock:2503; floor >=2.12 in pyproject.toml:30).  Its
published algorithm is restated in ``cistring.py``.
"""

from oracle import cistring, compound, contract, gates, givens, models, rand  # noqa: F401
