"""CPU oracle for the GRF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import anything from here, and
only as the checker or as the timed CPU baseline.  The product package
(``efficient-gaussian-process-on-graphs_b200/``) never imports ``oracle``: it
fails loudly when the CUDA library is missing.

Parity status
-------------
* sampler / step matrices / both ``fast_general_grf_kernel``s: PINNED.  The
  restatement in :mod:`oracle.grf_oracle` (numpy) and ``oracle/grf_oracle.c``
  (plain C) is checked bit-for-bit against outputs of the reference itself
  (imported from ``/root/reference`` in the build container by
  ``tests/golden/make_golden.py``; fixtures committed under ``tests/golden``).
* Phi(Phi^T V) matvec / CG / pathwise predict: parity UNPINNED against the
  reference (``gpytorch`` / ``linear_operator`` are not installed, the
  reference holds no golden vectors for this layer).  The oracle restates the
  reference's call sites with torch-CPU CSR ops and float64 scipy.
"""
