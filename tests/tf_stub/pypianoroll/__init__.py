"""TEST INFRASTRUCTURE ONLY -- placeholder for `pypianoroll` so that /root/reference/multinn/utils/data.py can be imported by
tools/make_golden_host.py (its NumPy functions pad_to_midi / load_data / prepare_sampling_inputs are what gets recorded;
the MIDI writer that needs the real package is not exercised)."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise NotImplementedError('pypianoroll is not installed; tests/tf_stub/pypianoroll is an import placeholder')


Multitrack = Track = _Unavailable
