"""TEST INFRASTRUCTURE ONLY -- `tensorflow.python.layers` of the NumPy TF stand-in."""
