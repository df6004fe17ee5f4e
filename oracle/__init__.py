"""CPU oracle for the PIGP hot path of ogaken1104/stopro.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or the timed CPU baseline.  ``stopro_b200`` never imports it.

Pinning status: **parity unpinned at the 1e-8 level by the reference's own
tests** -- the reference holds no golden vector or known-answer test for this
path and JAX is not installed here, so the reference cannot be executed.
What pins the oracle instead:

* ``sample_notebooks/sin_1D_direct.ipynb`` cell 17 prints the normalised
  initial loss 1.1445496082305908 (a float32 run); ``oracle.gp_ref`` gives
  1.14454947373... in float64 (tests/test_oracle.py::test_notebook_value).
* ``oracle.autodiff_ops`` rebuilds every differential operator exactly the way
  the reference does (nested grad / hessian of the scalar kernel, double
  vmap), with ``torch.func`` standing in for JAX; ``oracle.closed_form`` (the
  numpy restatement the GPU tests use at larger N) is checked against it.
* the reference tests' accuracy thresholds against analytic flow solutions.
"""
