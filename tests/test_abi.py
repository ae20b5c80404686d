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


def test_argument_errors_are_codes_not_crashes():
    """The boundary's error behaviour without a GPU: null pointers, non-positive sizes and unsupported shapes are
    rejected with MNN_ERR_ARG / MNN_ERR_UNSUPPORTED before anything touches the device, and the message is readable
    through mnn_last_error_string()."""
    import ctypes as C
    from multinn_b200 import _lib
    lib = _lib.lib
    ERR_ARG, ERR_UNSUPPORTED = -1, -2
    assert lib.mnn_gemm_tc(None, 4, 0, None, 4, 0, None, 4, None, 1.0, 0.0, 8, 8, 8, 0, None) == ERR_ARG
    assert 'null pointer' in _lib.last_error()
    buf = (C.c_float * 64)()
    p = C.cast(buf, C.c_void_p)
    assert lib.mnn_gemm_tc(p, 4, 0, p, 4, 0, p, 4, None, 1.0, 0.0, 0, 8, 8, 0, None) == ERR_ARG
    assert lib.mnn_gemm_tc(p, 3, 0, p, 4, 0, p, 4, None, 1.0, 0.0, 8, 8, 8, 0, None) == ERR_UNSUPPORTED   # TMA row stride
    assert lib.mnn_nade_logprob_fwd(None, p, 4, 0, 0, p, p, p, None, None, 0.0, 8, 1, 84, 256, 0, None) == ERR_ARG
    assert lib.mnn_nade_logprob_fwd(p, p, 4, 0, 0, p, p, p, None, None, 0.0, 8, 1, 84, 200, 0, None) == ERR_UNSUPPORTED
    assert 'num_hidden' in _lib.last_error()
    assert lib.mnn_nade_logprob_fwd(p, p, 4, 0, 0, p, p, p, None, None, 0.0, 8, 1, 84, 256, 4, None) == ERR_ARG   # stride < N
    assert lib.mnn_lstm_seq_fwd_tc(p, p, p, p, None, None, None, 1.0, 0, 4, 8, 12, p, 1, None) == ERR_UNSUPPORTED   # R % 8
    assert lib.mnn_lstm_tc_supported(8, 12) == 0 and lib.mnn_lstm_tc_supported(8, 16) == 1
    assert lib.mnn_colsum(None, 4, 4, 4, p, 0, p, None) == ERR_ARG
    assert lib.mnn_scale_rows(p, 4, 4, None, 4, 4, None) == ERR_ARG
    assert lib.mnn_set_sm_budget(-5) == 0 and lib.mnn_set_sm_budget(0) == 0
    # fused Gibbs chain: shapes it takes / refuses, and its argument checks
    assert lib.mnn_rbm_gibbs_smem_bytes(84, 256) == (2 * 84 * 256 + 8 * 4 * (84 + 256)) * 4
    assert lib.mnn_rbm_gibbs_smem_bytes(168, 84) > 0 and lib.mnn_rbm_gibbs_smem_bytes(84, 168) > 0
    assert lib.mnn_rbm_gibbs_smem_bytes(420, 168) == 0 and lib.mnn_rbm_gibbs_smem_bytes(84, 254) == 0
    assert lib.mnn_rbm_gibbs_smem_bytes(256, 256) == 0                      # 512 KB of weights
    g = lambda *a: lib.mnn_rbm_gibbs(*a)
    assert g(None, 84, p, None, 0, None, 0, None, None, 1, 0, 0, p, 84, p, 84, None, 0, 8, 84, 256, 2, None) == ERR_ARG
    assert g(p, 84, p, None, 0, None, 0, None, None, 0, 0, 0, p, 84, p, 84, None, 0, 8, 84, 256, 2, None) == ERR_ARG
    assert 'philox' in _lib.last_error()
    assert g(p, 84, p, None, 0, None, 0, p, None, 0, 0, 0, p, 84, p, 84, None, 0, 8, 84, 256, 2, None) == ERR_ARG   # uh without uv
    assert g(p, 420, p, None, 0, None, 0, None, None, 1, 0, 0, p, 420, p, 420, None, 0, 8, 420, 168, 2, None) == ERR_UNSUPPORTED
    assert g(p, 84, p, None, 0, None, 0, None, None, 1, 0, 0, p, 84, p, 84, None, 0, 8, 84, 256, 0, None) == ERR_ARG    # k == 0
    assert g(p, 84, p, None, 0, None, 0, None, None, 1, 0, 0, p, 83, p, 84, None, 0, 8, 84, 256, 2, None) == ERR_ARG    # ld_p % 4
    import pytest
    with pytest.raises(_lib.MultinnLibraryError):
        _lib.check(ERR_ARG, 'unit test')


def test_plain_c_program_compiles_against_the_header_and_links():
    """include/multinn_b200.h is the boundary: it must be valid C (not only C++), and a C program must link against the
    library and reach its argument checks. gcc -std=c99 -Wall -Werror; no GPU needed."""
    import shutil
    import subprocess
    import tempfile
    if shutil.which('gcc') is None:
        import pytest
        pytest.skip('gcc not available')
    import __graft_entry__ as g
    libdir = os.path.join(ROOT, 'multinn_b200')
    if not os.path.exists(os.path.join(libdir, 'libmultinn_sm100.so')):
        g.build()
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, 'abi_smoke')
        cuda_lib = '/usr/local/cuda/lib64'
        cmd = ['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'),
               os.path.join(ROOT, 'tests', 'abi_smoke.c'), '-o', exe, '-L', libdir, '-l:libmultinn_sm100.so',
               f'-Wl,-rpath,{libdir}', f'-Wl,-rpath-link,{cuda_lib}', '-Wl,--allow-shlib-undefined']
        subprocess.check_call(cmd)
        env = dict(os.environ, LD_LIBRARY_PATH=f'{libdir}:{cuda_lib}:' + os.environ.get('LD_LIBRARY_PATH', ''))
        out = subprocess.run([exe], env=env, capture_output=True, text=True)
        assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
        assert 'abi smoke ok' in out.stdout


def _header_prototypes():
    """{name: (return type, [parameter types])} parsed from the header's declarations."""
    src = open(os.path.join(ROOT, 'include', 'multinn_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'([A-Za-z_][\w\s\*]*?)\b(mnn_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;', src):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        types = []
        if params and params != 'void':
            for p in params.split(','):
                p = ' '.join(p.split())
                p = re.sub(r'\b[A-Za-z_]\w*$', '', p).strip() if not p.endswith('*') else p      # drop the parameter name
                types.append(p.replace(' *', '*'))
        protos[name] = (ret, types)
    return protos


def _ctype_of(ctype_str):
    import ctypes as C
    t = ctype_str.replace('const ', '').strip()
    if t.endswith('*') or t == 'mnn_stream_t':
        return C.c_char_p if ctype_str.strip() == 'const char*' else C.c_void_p
    return {'int': C.c_int, 'long long': C.c_longlong, 'float': C.c_float, 'unsigned long long': C.c_ulonglong,
            'size_t': C.c_size_t, 'uint32_t': C.c_uint32}[t]


def test_ctypes_signatures_match_the_header_prototypes():
    """Every argtypes / restype entry of multinn_b200/_lib.py against the C prototype in the header: arity, integer widths
    (int vs long long vs size_t), floats, pointers. A mismatch here corrupts arguments silently at call time."""
    import ctypes as C
    from multinn_b200 import _lib
    protos = _header_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    for name, (ret, params) in protos.items():
        want = [_ctype_of(p) for p in params]
        got = list(_lib.SIGNATURES[name])
        assert len(got) == len(want), f'{name}: {len(got)} ctypes arguments for {len(want)} C parameters'
        for i, (g, w) in enumerate(zip(got, want)):
            assert g is w, f'{name}: argument {i} is {g.__name__} in _lib.py but `{params[i]}` in the header'
        want_ret = _ctype_of(ret)
        got_ret = _lib._RESTYPES.get(name, C.c_int)
        assert got_ret is want_ret, f'{name}: restype {got_ret.__name__} but the header returns `{ret}`'
