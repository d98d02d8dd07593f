"""CPU oracle for the wavelet hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product path (``wavelet_transformer_b200``) never imports this
package and fails loudly when the CUDA library is missing.

Parity status (see DESIGN.md "Oracle"):

* ``oracle.pycwt_oracle`` / ``oracle.pywt_oracle`` restate pycwt 0.4.0b0 and
  PyWavelets 1.9.0 (pinned by /root/reference/requirements.txt:33,39).  Neither
  package is vendored in the reference nor installable here, and the
  reference's tests hold no golden vectors for them: **parity unpinned** at
  those two boundaries.  The restatement is anchored on the reference's call
  sites and on analytic / algebraic known-answer tests (tests/test_oracle_*.py).
* ``oracle.modwt_oracle`` is pinned: it is checked against the reference's own
  ``src/modwt.py`` arithmetic, executed in the build container, through the
  committed fixtures in ``tests/golden/`` (made by tests/golden/make_golden.py).
"""
