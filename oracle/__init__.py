"""CPU oracle for the cuda_b200 hot path — TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the *reference numpy backend* of
google/TensorNetwork (v0.4.6) for the rows of SURVEY.md section 8(a).  It exists so that
the CUDA path can be checked for parity on the GPU box, where `/root/reference`
does not exist.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import it.  The product package
`tensornetwork_b200` never imports it and has no CPU fallback.

Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py)
against golden vectors produced by running the real reference (imported from
/root/reference through a 3-module import shim) in the build container; the
generating script is `oracle/gen_golden.py`, the vectors live in `tests/golden/`.
The known-answer tests of the reference's own suites are restated in
tests/test_oracle_kat.py (file:line cited per test).
"""
