"""CPU dry run of the ops wrappers: the C library is replaced by a recorder that checks every call against the ctypes
signature table (arity, ints where the ABI takes ints, floats where it takes floats, pointers or None elsewhere), and
the CUDA-only plumbing (device check, current stream) is stubbed. Catches drift between multinn_b200/ops.py and the ABI
without a GPU; numerical behaviour is the GPU tests' business."""
import ctypes as C

import pytest
import torch

from multinn_b200 import _lib, ops


class _Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        sig = _lib.SIGNATURES[name]            # KeyError = wrapper calls a symbol the table does not know

        def fn(*args):
            assert len(args) == len(sig), f'{name}: {len(args)} arguments for {len(sig)} parameters'
            for i, (a, t) in enumerate(zip(args, sig)):
                if t in (C.c_int, C.c_longlong, C.c_ulonglong, C.c_size_t):
                    assert isinstance(a, int) and not isinstance(a, bool), f'{name}: argument {i} = {a!r} is not an int'
                elif t is C.c_float:
                    assert isinstance(a, float), f'{name}: argument {i} = {a!r} is not a float'
                else:
                    assert a is None or isinstance(a, int), f'{name}: argument {i} = {a!r} is not a pointer'
            self.calls.append(name)
            if name.endswith('_supported'):
                return 1
            if name.endswith('_bytes'):
                return 1024
            if name.endswith('_ctas'):
                return 8
            return 0
        return fn


@pytest.fixture
def dry(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(ops, 'lib', rec)
    monkeypatch.setattr(ops, '_ptr', lambda t: None if t is None else t.data_ptr())
    monkeypatch.setattr(ops, '_stream', lambda: 0)

    class _S:
        cuda_stream = 0
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: _S())
    return rec


def test_every_wrapper_matches_the_abi_signature(dry):
    f = lambda *s: torch.zeros(*s)
    B, T, D, M, H, R, N = 2, 3, 84, 5, 128, 16, 6
    x = f(B, T, D, M)
    ops.pack_pianoroll(x, f(T + 1, B, D * M), f(M, T + 1, B, D), torch.zeros(M, T * B, 4, dtype=torch.int32))
    ops.pack_pianoroll(x.to(torch.uint8), f(T + 1, B, D * M))
    ops.pack_rows(f(N, D), torch.zeros(N, 4, dtype=torch.int32))
    ops.gemm(f(4, 8), f(8, 12), f(4, 12), bias=f(12), alpha=2, beta=1)
    ops.gemm(f(8, 4), f(12, 8), f(4, 12), transA=True, transB=True, mode='f32')
    ops.set_sm_budget(32)
    with ops.gemm_split('pair'):
        ops.gemm(f(4, 8), f(8, 12), f(4, 12))
        ops.gemm(f(256, 64), f(64, 136), f(256, 136), b_weight=True)       # pre-split weight -> mnn_gemm_tc_bpair
        xb = f(2, 255, 16, 4)                                             # [B,T,D,M]: 512 stacked rows of 64 features
        base, twin = f(512, 64), torch.zeros(512, 64, dtype=torch.int16)
        ops.pack_stacked_bf16(xb, twin)
        ops.register_twin(base, twin)
        ops.gemm(base[256:], f(64, 136), f(256, 136), a_exact=True, b_weight=True)      # bf16 A plane -> mnn_gemm_tc_abf16
        ops.gemm(base, f(512, 136), f(64, 136), transA=True, a_exact=True)              # (too small for the pair kernel)
    assert ops.lstm_seq_ctas(T, B, R, 64) == 8 and ops.lstm_seq_ctas(T, B, R, 64, backward=True) == 8
    ops.colsum(f(N, 12), f(12), accumulate=True)
    ops.lstm_cell_fwd(f(B, 4 * R), f(B, R), f(B, R), f(B, R), out=f(B, R), dscale=f(B, R), u=f(B, R), keep=0.9, seed=3)
    gates, wh, hb, cb = f(T, B, 4 * R), f(R, 4 * R), f(T + 1, B, R), f(T + 1, B, R)
    for mode in ('tc', 'simt'):
        ops.lstm_seq_fwd(gates, wh, hb, cb, out=f(T, B, R), dscale=f(T, B, R), keep=0.9, seed=5, mode=mode)
        ops.lstm_seq_bwd(gates, wh, cb, f(T, B, R), f(T, B, R), f(B, R), f(B, R), mode=mode)
    ops.lstm_seq_bwd(gates, wh, cb, f(T, B, R), f(T, B, R), f(B, R), f(B, R), has_next=True)
    bits = torch.zeros(M, N, 4, dtype=torch.int32)
    fc, we, wd = f(N, M * (H + D)), f(M, D, H), f(M, D, H)
    ops.nade_logprob_fwd(bits, fc, 0, M * H, we, wd, f(M, N), cond_p=f(M, N, D), dfc=torch.zeros_like(fc), gscale=0.5)
    ops.nade_logprob_bwd(bits, fc, 0, M * H, we, wd, torch.zeros_like(fc), torch.zeros_like(we), torch.zeros_like(wd))
    out = f(N, M * D)
    ops.nade_sample(fc, 0, M * H, we, wd, out, out.stride(0), M, 1, u=f(M, N, D), nll=f(M, N))
    ops.nade_sample(fc, 0, M * H, we, wd, out, out.stride(0), M, 1, use_philox=True, seed=7, offset=2)
    ops.bias_sigmoid_sample(f(N, H), bias=f(N, H), u=f(N, H), p=f(N, H), s=f(N, H))
    ops.bias_sigmoid_sample(f(N, H), bias=f(1, H), s=f(N, H), use_philox=True, seed=1, offset=9)
    assert ops.rbm_gibbs_supported(f(N, D), f(D, H), f(N, H), f(1, D), (f(2, N, H), f(2, N, D)))
    ops.rbm_gibbs(f(N, D), f(D, H), f(N, H), f(1, D), 2, p_v=f(N, D), v_k=f(N, D), h_k=f(N, H), u=(f(2, N, H), f(2, N, D)))
    ops.rbm_gibbs(f(N, D), f(D, H), None, None, 2, v_k=f(N, D), seed=4, offset=8)
    ops.sigmoid_bwd(f(N, H), f(N, H), f(N, H))
    ops.rbm_free_energy(f(N, H), f(1, H), f(N, D), f(N, D), f(N))
    ops.sum_into(f(100), f(1), scale=0.5, accumulate=True)
    ops.sqnorm_into(f(100), f(1))
    ops.clip_adam(f(100), f(100), f(100), f(100), f(1), 3, 0.01)
    ops.clip_sgd(f(100), f(100), f(1), 0.01)
    ops.scale_rows(f(N, 12), f(N))
    ops.axpy(f(100), f(100), -0.5)
    Rg = (16, 8)
    kern = [f(M * D + Rg[0], 4 * Rg[0]), f(Rg[0] + Rg[1], 4 * Rg[1])]
    assert ops.generate_fused_supported(2, M * D, Rg, B, M, D, H)
    ops.generate_fused(kern, [f(4 * Rg[0]), f(4 * Rg[1])], [(f(B, Rg[0]), f(B, Rg[0])), (f(B, Rg[1]), f(B, Rg[1]))],
                       f(Rg[1], M * (H + D)), f(M * (H + D)), we, wd, f(B, M * (H + D)), f(B, 3, M * D), seed=5)
    check_probe = ops.lib.mnn_probe_mufu(None, 8, 128, 4, 0, 0)
    assert check_probe == 0
    ops.set_nade_mode('simt')
    ops.set_nade_mode('tc')
    with ops.row_map(B, 4 * B, B):
        assert ops.row_scale() == 4
        with ops.row_map_scaled(3):
            assert ops.row_scale() == 4
        ops.lstm_seq_fwd(gates, wh, hb, cb, out=f(T, B, R), dscale=f(T, B, R), keep=0.9, seed=5, t_base=32)
    assert ops.row_scale() == 1
    called = set(dry.calls)
    never = set(_lib.SIGNATURES) - called - {'mnn_version', 'mnn_last_error_string', 'mnn_launch_count',
                                             'mnn_lstm_seq_bwd_tc'}      # superseded by the _chunk entry point
    assert not never, f'ABI entry points no wrapper reaches: {sorted(never)}'
