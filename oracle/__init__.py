"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the MultINN hot path.

PARITY UNPINNED: the reference (ilya16/MultINN) ships no tests, golden vectors or
fixtures for this path, and its arithmetic lives in un-vendored third-party wheels
(tensorflow==1.13.1, tensorflow_probability==0.6.0, reference requirements.txt:4-6)
that cannot be installed on Python 3.12 / offline. The oracle therefore pins itself
(two independent restatements + analytic known-answer tests, see tests/test_oracle_*.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package. Nothing under `multinn_b200/` may import it.
"""
