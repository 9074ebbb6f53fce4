"""CPU oracle for the per-step pseudo-spectral update (TEST INFRASTRUCTURE ONLY).

This package is a numpy restatement of the algorithm behind the reference's
``Operator.integrate(u_0, mesh, dt, step)`` hot path (qiauil/torchfsm v0.0.4).
It exists to CHECK the CUDA path; it is never the thing shipped or measured:

* only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
  ``--impl reference`` legs of ``bench.py`` may import it;
* nothing under ``torchfsm_b200/`` imports it (tests/test_product_boundary.py::test_product_does_not_import_oracle_or_emulator
  enforces that), and the product fails loudly when the CUDA library is missing.

Arithmetic lives in a third-party dependency of the reference that is not under
``/root/reference``: PyTorch (``requirements.txt:4``, unpinned; the version in
force when the golden vectors were generated was torch 2.11.0+cu128, CPU/MKL).
The oracle restates the reference's call sites (file:line cited on every
function) over numpy + ``scipy.fft`` (pocketfft), keeping the reference's full
complex (C2C) spectrum layout ``(B, C, N...)`` and its evaluation order.

Parity pinning: the reference ships no tests, goldens or fixtures,
so the oracle is pinned against outputs of the reference itself, generated in the
authoring container by ``tests/golden/make_golden.py`` (which imports
``/root/reference`` read-only) and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every fixture (fp64 <= 1e-12, fp32 <= 1e-5
relative L2) plus analytic known answers.
"""

from .spectral import OracleMesh  # noqa: F401
from .operator import OracleOperator  # noqa: F401
