"""TEST INFRASTRUCTURE ONLY -- `tensorflow.python` package of the NumPy TF stand-in (see tests/tf_stub/tensorflow/__init__.py)."""
