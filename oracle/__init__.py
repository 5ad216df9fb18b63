"""oracle/ -- TEST INFRASTRUCTURE, not product code.

A CPU restatement (numpy / scipy / plain Python) of the tracking-by-detection hot path of
AdaptiveCity/deepdish, used ONLY as the checker for the CUDA path:

  * ``tests/``                      -- parity tests (CUDA vs oracle, oracle vs golden fixtures)
  * ``__graft_entry__.smoke()``     -- one small GPU invocation checked against the oracle
  * ``bench.py``                    -- the ``cpu_baseline`` leg and ``--impl reference`` arm

Nothing under ``deepdish_b200/`` imports, links or executes anything in this package; the product
path raises if the CUDA library is missing instead of falling back to this code.

Parity pinning.  The reference has no test-suite; its only known-answer vectors are the six
asserts in ``tools/intersection.py:35-57`` (replayed in ``tests/test_oracle_intersection.py``).
Everything else is pinned by running the *unmodified* reference (imported from ``/root/reference``
with three shims, see ``oracle/refload.py``) in the build container and committing its outputs as
fixtures under ``tests/golden/`` together with the generating script ``oracle/make_golden.py``.
The one piece with no reference-side pin is the SSD 1917-anchor decode (third-party TFLite op
``TFLite_Detection_PostProcess``, absent from ``/root/reference``): ``oracle/detect.py`` restates
its published algorithm and says "parity unpinned" there.
"""
