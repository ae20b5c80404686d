"""CPU: the C-ABI library loads and exports every symbol include/multinn_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'multinn_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mnn_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, 'multinn_b200', 'libmultinn_sm100.so')):
        g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, 'multinn_b200', 'libmultinn_sm100.so'))
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'
    lib.mnn_version.restype = ctypes.c_int
    assert lib.mnn_version() >= 100


def test_binding_table_matches_header():
    from multinn_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    for n in _lib.SIGNATURES:
        assert getattr(_lib.lib, n).argtypes is not None


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: ops raise on host tensors instead of computing elsewhere."""
    import pytest
    import torch
    from multinn_b200 import ops
    with pytest.raises(ValueError):
        ops.gemm(torch.zeros(2, 2), torch.zeros(2, 2), torch.zeros(2, 2))


def test_product_does_not_import_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, 'multinn_b200')):
        for f in fs:
            if f.endswith('.py'):
                s = open(os.path.join(dp, f)).read()
                assert 'oracle' not in s.replace('# oracle', ''), f'{f} mentions the oracle'
