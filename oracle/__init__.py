"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's decode + NMS path.

Nothing in ``pytorch_yolo_b200/`` (the product) may import this package.  The only
permitted users are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker / the
timed CPU baseline, never as the thing shipped.

Parity pin: the reference (Dipet/pytorch_yolo) ships no tests and no golden
vectors, so the pin is the reference itself, run in the build
container: ``tests/golden/make_golden.py`` imports the unmodified reference
functions by path, runs them on seeded inputs and commits inputs + outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against
those vectors bit-for-bit, and (when ``/root/reference`` is present) against the
live reference.
"""
