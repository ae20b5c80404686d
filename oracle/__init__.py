"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the MultINN hot path.

PARITY PINNED BY THE REFERENCE'S OWN CODE ON A NUMPY STAND-IN FOR TENSORFLOW (not by TensorFlow itself): the reference
(ilya16/MultINN) ships no tests, golden vectors or fixtures for this path, and its arithmetic lives in un-vendored
third-party wheels (tensorflow==1.13.1, tensorflow_probability==0.6.0, reference requirements.txt:4-6) that cannot be
installed on Python 3.12 / offline. tools/make_golden_ref.py therefore imports the reference's modules UNMODIFIED
(models/common/{nade,rbm,dbn}.py, utils/sequences.py, metrics/statistical.py, generators/rnn_multinade.py) on top of
tests/tf_stub/ (NumPy-backed `tensorflow` / `tensorflow_probability`) and commits their outputs as
tests/golden/ref_primitives.npz; tests/test_ref_golden.py checks this oracle (and the CUDA path) against them. Loop
structure, transposes, eps placement, bias split, Gibbs / CD-k structure and flatten order are pinned by the reference's
code; the semantics of the individual ops are NumPy's. The LSTM cell, tf.gradients and the optimiser
are pinned by torch.nn.LSTM, torch autograd / finite differences and closed forms (tests/test_oracle_kat.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package. Nothing under `multinn_b200/` may import it.
"""
