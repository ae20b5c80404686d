"""Thin torch-tensor wrappers over the C ABI (include/multinn_b200.h). torch is plumbing only: device
memory, the current CUDA stream and torch.distributed. Every op launches on torch's current stream.
"""
import contextlib
import os

import torch

from ._lib import check, lib

GEMM_MODE = "tc"   # "tc": tcgen05 3xTF32 GEMM (fp32-accurate); "f32": CUDA-core fp32 GEMM (checker / unaligned operands)


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("multinn_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype not in (torch.float32, torch.int32, torch.uint8, torch.bool, torch.float64):
        raise ValueError(f"unsupported dtype {t.dtype}")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rowstride(t):
    assert t.dim() == 2 and t.stride(1) == 1, "need a row-major 2-D view"
    return t.stride(0)


def pack_pianoroll(x, xin=None, xtr=None, bits=None):
    """x[B,T,D,M] -> xin[T+1,B,D*M], xtr[M,T+1,B,D], bits[M,T*B,4] (any subset)."""
    B, T, D, M = x.shape
    assert x.is_contiguous()
    fn = lib.mnn_pack_pianoroll if x.dtype == torch.float32 else lib.mnn_pack_pianoroll_u8
    check(fn(_ptr(x), _ptr(xin), _ptr(xtr), _ptr(bits), B, T, D, M, _stream()), "pack_pianoroll")


def pack_rows(v, bits, D=None, dim_stride=1):
    """bits[N,4] of a binary matrix view v[N, ...] (element (n,d) at v[n*ld + d*dim_stride])."""
    N = v.shape[0]
    ld = v.stride(0)
    D = D if D is not None else v.shape[1]
    check(lib.mnn_pack_rows(_ptr(v), ld, dim_stride, _ptr(bits), N, D, _stream()), "pack_rows")


_gemm_split = '2.5'
_pair_cache = {}          # weights pre-split to bf16 pairs, valid inside one gemm_split('pair') block (= one training step)
PRESPLIT_WEIGHTS = os.environ.get('MNN_PRESPLIT_WEIGHTS', '1') != '0'


def split_bf16_pair(W):
    """bf16 pair planes [2, rows, ld] (int16 storage) of a row-major fp32 matrix view; ld = cols rounded up to 8."""
    rows, cols = W.shape
    ldp = (cols + 7) // 8 * 8
    out = torch.empty(2, rows, ldp, dtype=torch.int16, device=W.device)
    check(lib.mnn_split_bf16_pair(_ptr(W), _rowstride(W), rows, cols, out.data_ptr(), ldp, _stream()), "split_bf16_pair")
    return out


_twins = []               # (fp32 base tensor, its exact bf16 plane [rows, ld16]) registered by the input staging
BF16_INPUT_TWIN = os.environ.get('MNN_BF16_INPUT_TWIN', '1') != '0'


def pack_stacked_bf16(x, out):
    """x[B,T,D,M] (float32 / uint8 / bool) -> out[(T+1)*B, ld16] int16 storage of the exact bf16 stacked rows."""
    B, T = x.shape[0], x.shape[1]
    I = x.shape[2] * x.shape[3]
    assert x.is_contiguous() and out.is_contiguous() and out.shape[0] == (T + 1) * B
    check(lib.mnn_pack_stacked_bf16(_ptr(x), int(x.dtype != torch.float32), out.data_ptr(), out.shape[1], B, T, I, _stream()),
          "pack_stacked_bf16")


def register_twin(base, twin):
    """`twin` holds the same rows as the fp32 matrix `base` [rows, cols] as exact bf16 (binary data): GEMMs of the training
    step whose A operand is a row block of `base` (a_exact=True) read the twin instead. One registration per base."""
    _twins[:] = [(b, t) for b, t in _twins if b.data_ptr() != base.data_ptr()]
    _twins.append((base, twin))
    del _twins[:-4]


def _twin_of(A):
    """(pointer, row stride in elements) of the bf16 rows matching the row-block view A of a registered base, or None."""
    for base, twin in _twins:
        off = A.data_ptr() - base.data_ptr()
        row_bytes = base.stride(0) * 4
        if 0 <= off < base.shape[0] * row_bytes and off % row_bytes == 0 and A.stride(0) == base.stride(0) \
                and A.shape[1] == base.shape[1] and off // row_bytes + A.shape[0] <= base.shape[0]:
            return twin.data_ptr() + (off // row_bytes) * twin.shape[1] * 2, twin.shape[1]
    return None


def gemm(A, B, C, transA=False, transB=False, bias=None, alpha=1.0, beta=0.0, a_exact=False, mode=None, b_weight=False):
    """C[M,N] = alpha * op(A) op(B) + beta * C (+ bias). A, B, C are row-major 2-D views (row stride free).
    mode 'tc': tcgen05 3xTF32 (fp32-accurate) kernel; 'f32': CUDA-core fp32 kernel (also used when an operand is
    not TMA-addressable). a_exact: A holds only tf32-exact values (binary inputs) -> 2 products instead of 3.
    b_weight: B is a weight matrix that does not change inside the enclosing gemm_split('pair') block: it is split to
    bf16 pairs once per block (cached by view) and handed to the kernel pre-split."""
    M, N = C.shape
    K = A.shape[0] if transA else A.shape[1]
    assert (A.shape[1] if transA else A.shape[0]) == M
    assert (B.shape[1] if transB else B.shape[0]) == K and (B.shape[0] if transB else B.shape[1]) == N
    mode = mode or GEMM_MODE
    pair_ok = _gemm_split == 'pair' and mode == 'tc' and M >= 256 and N > 128 and K >= 64
    if a_exact and pair_ok and BF16_INPUT_TWIN and _twins:
        tw = _twin_of(A)
        if tw is not None:
            bp = None
            if b_weight and PRESPLIT_WEIGHTS:
                key = (B.data_ptr(), tuple(B.shape), B.stride(0), torch.cuda.current_stream().cuda_stream)
                bp = _pair_cache.get(key)
                if bp is None:
                    bp = _pair_cache[key] = split_bf16_pair(B)
            elif _rowstride(B) % 4 or B.data_ptr() % 16:
                tw = None
            if tw is not None:
                check(lib.mnn_gemm_tc_abf16(tw[0], tw[1], int(transA), _ptr(B), _rowstride(B),
                                            None if bp is None else bp.data_ptr(), 0 if bp is None else bp.shape[2],
                                            int(transB), _ptr(C), _rowstride(C), _ptr(bias), float(alpha), float(beta),
                                            M, N, K, _stream()), "gemm_tc_abf16")
                return
    if (b_weight and PRESPLIT_WEIGHTS and _gemm_split == 'pair' and mode == 'tc' and M >= 256 and N > 128 and K >= 64
            and _rowstride(A) % 4 == 0 and A.data_ptr() % 16 == 0):
        key = (B.data_ptr(), tuple(B.shape), B.stride(0), torch.cuda.current_stream().cuda_stream)
        pair = _pair_cache.get(key)
        if pair is None:
            pair = _pair_cache[key] = split_bf16_pair(B)
        check(lib.mnn_gemm_tc_bpair(_ptr(A), _rowstride(A), int(transA), pair.data_ptr(), pair.shape[2], int(transB),
                                    _ptr(C), _rowstride(C), _ptr(bias), float(alpha), float(beta), M, N, K, int(a_exact),
                                    _stream()), "gemm_tc_bpair")
        return
    if mode == 'tc' and lib.mnn_gemm_tc_supported(_ptr(A), _rowstride(A), _ptr(B), _rowstride(B)):
        check(lib.mnn_gemm_tc(_ptr(A), _rowstride(A), int(transA), _ptr(B), _rowstride(B), int(transB), _ptr(C),
                              _rowstride(C), _ptr(bias), float(alpha), float(beta), M, N, K, int(a_exact), _stream()),
              "gemm_tc")
        return
    check(lib.mnn_gemm_f32(_ptr(A), _rowstride(A), int(transA), _ptr(B), _rowstride(B), int(transB), _ptr(C),
                           _rowstride(C), _ptr(bias), float(alpha), float(beta), M, N, K, _stream()), "gemm_f32")


def set_gemm_split(mode):
    """Operand split of the pair GEMM for this thread's following launches: '2.5' (default: tf32 main product + bf16 cross
    terms, ~2^-20 per product) or 'pair' (bf16 pairs, three kind::f16 MMAs, ~2^-17, 9 % faster; the training step)."""
    check(lib.mnn_set_gemm_split({'2.5': 0, 'pair': 1}[mode]), "set_gemm_split")


@contextlib.contextmanager
def gemm_split(mode):
    global _gemm_split
    set_gemm_split(mode)
    _gemm_split = mode
    _pair_cache.clear()
    del _twins[:]
    try:
        yield
    finally:
        set_gemm_split('2.5')
        _gemm_split = '2.5'
        _pair_cache.clear()
        del _twins[:]


def set_sm_budget(sms):
    """Cap the persistent GEMM grids launched by this thread at `sms` SMs (0 = all)."""
    check(lib.mnn_set_sm_budget(int(sms)), "set_sm_budget")


def lstm_seq_ctas(T, B, R, budget, backward=False):
    """SMs the persistent recurrence kernel of one layer occupies when launched under `budget`."""
    set_sm_budget(budget)
    try:
        fn = lib.mnn_lstm_seq_bwd_ctas if backward else lib.mnn_lstm_seq_fwd_ctas
        return int(fn(int(T), int(B), int(R)))
    finally:
        set_sm_budget(0)


def num_sms():
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


_colsum_ws = {}


def colsum(A, out, accumulate=False):
    rows, cols = A.shape
    key = (A.device.index, torch.cuda.current_stream().cuda_stream)
    need = int(lib.mnn_colsum_workspace_bytes(cols))
    ws = _colsum_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = _colsum_ws[key] = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=A.device)
    check(lib.mnn_colsum(_ptr(A), _rowstride(A), rows, cols, _ptr(out), int(accumulate), _ptr(ws), _stream()),
          "colsum")


# ----------------------------------------------------------------------------- data-parallel noise keying
_row_map = (0, 0, 0)


class row_map:
    """Context manager around launches that may draw Philox noise: local row r of the calling rank is keyed as global
    row (r // rows_local) * rows_global + row_base + r % rows_local (mnn_set_row_map), so dropout masks, Bernoulli codes
    and samples do not depend on how the batch is sharded over GPUs. Nests; restores the previous map on exit."""

    def __init__(self, rows_local, rows_global, row_base):
        self.new = (int(rows_local), int(rows_global), int(row_base))

    def __enter__(self):
        global _row_map
        self.old, _row_map = _row_map, self.new
        check(lib.mnn_set_row_map(*self.new), "set_row_map")
        return self

    def __exit__(self, *exc):
        global _row_map
        _row_map = self.old
        check(lib.mnn_set_row_map(*self.old), "set_row_map")
        return False


def row_map_scaled(k):
    """The active map with every size multiplied by k: for b-major [B*k, .] row blocks (k generated steps per sequence)."""
    return row_map(_row_map[0] * k, _row_map[1] * k, _row_map[2] * k)


def row_scale():
    """rows_global / rows_local of the active row map (1 without one): per-call Philox offsets advance by GLOBAL counts."""
    return _row_map[1] // _row_map[0] if _row_map[0] else 1


def lstm_cell_fwd(gates, c_prev, c, h, out=None, dscale=None, u=None, keep=1.0, seed=0, offset=0):
    B, R4 = gates.shape
    check(lib.mnn_lstm_cell_fwd(_ptr(gates), _ptr(c_prev), _ptr(c), _ptr(h), _ptr(out), _ptr(dscale), _ptr(u),
                                float(keep), seed, offset, B, R4 // 4, _stream()), "lstm_cell_fwd")


LSTM_MODE = "tc"          # "tc": tcgen05 recurrence with fused cell; "simt": per-step fp32 GEMM + cell kernels
LSTM_PERSISTENT = True    # one cooperative launch for all T steps (tc mode)
_lstm_ws = {}


def _lstm_workspace(B, R, device):
    key = (device.index, torch.cuda.current_stream().cuda_stream, B, R)
    ws = _lstm_ws.get(key)
    if ws is None:
        ws = _lstm_ws[key] = torch.empty(int(lib.mnn_lstm_workspace_bytes(B, R)), dtype=torch.uint8, device=device)
    return ws


def lstm_seq_fwd(gates, wh, hbuf, cbuf, out=None, dscale=None, u=None, keep=1.0, seed=0, mode=None, persistent=None,
                 t_base=0):
    """t_base: index of gates[0]'s time step inside the whole sequence (chunked launches): the Philox dropout counter is
    (global batch row, unit, t_base + t), so every chunking, kernel variant and GPU count draws the same masks."""
    T, B, R4 = gates.shape
    assert gates.is_contiguous() and hbuf.is_contiguous() and cbuf.is_contiguous()
    assert wh.stride(1) == 1 and wh.stride(0) == R4
    mode = mode or LSTM_MODE
    if t_base:
        check(lib.mnn_set_time_base(int(t_base)), "set_time_base")
    try:
        if mode == "tc" and lib.mnn_lstm_tc_supported(B, R4 // 4):
            pers = LSTM_PERSISTENT if persistent is None else persistent
            check(lib.mnn_lstm_seq_fwd_tc(_ptr(gates), _ptr(wh), _ptr(hbuf), _ptr(cbuf), _ptr(out), _ptr(dscale), _ptr(u),
                                          float(keep), seed, T, B, R4 // 4,
                                          _ptr(_lstm_workspace(B, R4 // 4, gates.device)), int(pers), _stream()),
                  "lstm_seq_fwd_tc")
            return
        check(lib.mnn_lstm_seq_fwd(_ptr(gates), _ptr(wh), _ptr(hbuf), _ptr(cbuf), _ptr(out), _ptr(dscale), _ptr(u),
                                   float(keep), seed, T, B, R4 // 4, _stream()), "lstm_seq_fwd")
    finally:
        if t_base:
            check(lib.mnn_set_time_base(0), "set_time_base")


def lstm_seq_bwd(gates, wh, cbuf, dout, dscale, dh_work, dc_work, mode=None, persistent=None, has_next=False):
    """has_next: `gates` is a time chunk of a longer buffer whose following chunk was already back-propagated."""
    T, B, R4 = gates.shape
    mode = mode or LSTM_MODE
    if mode == "tc" and lib.mnn_lstm_tc_supported(B, R4 // 4):
        pers = LSTM_PERSISTENT if persistent is None else persistent
        check(lib.mnn_lstm_seq_bwd_tc_chunk(_ptr(gates), _ptr(wh), _ptr(cbuf), _ptr(dout), _ptr(dscale), _ptr(dc_work),
                                            T, B, R4 // 4, _ptr(_lstm_workspace(B, R4 // 4, gates.device)), int(pers),
                                            int(has_next), _stream()), "lstm_seq_bwd_tc")
        return
    if has_next:
        raise ValueError("chunked BPTT needs the tensor-core recurrence (num_units % 8 == 0)")
    check(lib.mnn_lstm_seq_bwd(_ptr(gates), _ptr(wh), _ptr(cbuf), _ptr(dout), _ptr(dscale), _ptr(dh_work),
                               _ptr(dc_work), T, B, R4 // 4, _stream()), "lstm_seq_bwd")


def set_nade_mode(mode):
    """'simt' (default): the CUDA-core segment kernel; 'tc': the tcgen05 segment-row forward kernel where the shape fits
    (parity-tested, slower today: DESIGN.md section 9)."""
    check(lib.mnn_set_nade_mode({'tc': 0, 'simt': 1}[mode]), "set_nade_mode")


def nade_logprob_fwd(bits, fc, enc_col0, dec_col0, w_enc, w_dec, nll, cond_p=None, dfc=None, gscale=0.0):
    """bits[M,N,4], nll[M,N], cond_p[M,N,D] may be row slices ([:, r0:r1]) of longer buffers; fc/dfc the same rows."""
    M, D, H = w_enc.shape
    N = fc.shape[0]
    ts = bits.stride(0) // bits.stride(1) if M > 1 else N
    assert bits.shape[1] == N and nll.shape[1] == N and nll.stride(1) == 1 and (M == 1 or nll.stride(0) == ts)
    assert cond_p is None or M == 1 or (cond_p.stride(0) == ts * D and cond_p.stride(1) == D)
    assert dfc is None or dfc.stride(0) == fc.stride(0)
    check(lib.mnn_nade_logprob_fwd(_ptr(bits), _ptr(fc), _rowstride(fc), enc_col0, dec_col0, _ptr(w_enc), _ptr(w_dec),
                                   _ptr(nll), _ptr(cond_p), _ptr(dfc), float(gscale), N, M, D, H, ts, _stream()),
          "nade_logprob_fwd")


def nade_logprob_bwd(bits, fc, enc_col0, dec_col0, w_enc, w_dec, dfc, dw_enc, dw_dec):
    M, D, H = w_enc.shape
    N = fc.shape[0]
    ts = bits.stride(0) // bits.stride(1) if M > 1 else N
    assert bits.shape[1] == N and dfc.stride(0) == fc.stride(0)
    check(lib.mnn_nade_logprob_bwd(_ptr(bits), _ptr(fc), _rowstride(fc), enc_col0, dec_col0, _ptr(w_enc), _ptr(w_dec),
                                   _ptr(dfc), _ptr(dw_enc), _ptr(dw_dec), N, M, D, H, ts, _stream()), "nade_logprob_bwd")


def nade_sample(fc, enc_col0, dec_col0, w_enc, w_dec, out, out_ld, out_dim_stride, out_track_stride, u=None,
                use_philox=False, seed=0, offset=0, nll=None):
    M, D, H = w_enc.shape
    N = fc.shape[0]
    check(lib.mnn_nade_sample(_ptr(fc), _rowstride(fc), enc_col0, dec_col0, _ptr(w_enc), _ptr(w_dec), _ptr(u),
                              int(use_philox), seed, offset, _ptr(out), out_ld, out_dim_stride, out_track_stride,
                              _ptr(nll), N, M, D, H, _stream()), "nade_sample")


def bias_sigmoid_sample(pre, bias=None, u=None, p=None, s=None, use_philox=False, seed=0, offset=0):
    """p = sigmoid(pre + bias); s = float(u < p). bias is [C]/[1,C] (broadcast) or [N,C] (per row)."""
    N, Cc = pre.shape
    ld_bias = 0
    if bias is not None and bias.dim() == 2 and bias.shape[0] == N and N > 1:
        ld_bias = bias.stride(0)
    check(lib.mnn_bias_sigmoid_sample(_ptr(pre), _rowstride(pre), _ptr(bias), ld_bias, _ptr(u),
                                      _rowstride(u) if u is not None else 0, int(use_philox), seed, offset,
                                      _ptr(p), _rowstride(p) if p is not None else 0, _ptr(s),
                                      _rowstride(s) if s is not None else 0, N, Cc, _stream()), "bias_sigmoid_sample")


# k-step Gibbs chain: "fused" = one launch of mnn_rbm_gibbs where the shape fits; "gemm" = 2k tensor-core GEMM + half-step
# launches; "auto" picks by the measured crossover (profiles/r1_gibbs_bench.log, 84 x 256, k = 10 on a B200: fused 0.21 ms
# per wave of 4 736 rows, GEMM path 0.67 ms of launches + 0.026 ms per 1 000 rows -> fused wins below ~37 k rows: 3.1x at
# the generation batch sizes, 0.83x at C3's 65 536 training rows).
GIBBS_MODE = "auto"
GIBBS_FUSED_MAX_ROWS = 32768


def gibbs_use_fused(n_rows):
    return GIBBS_MODE == "fused" or (GIBBS_MODE == "auto" and n_rows <= GIBBS_FUSED_MAX_ROWS)


def _bias_ld(b, N):
    if b is None:
        return 0
    return b.stride(0) if (b.dim() == 2 and b.shape[0] == N and N > 1) else 0


def rbm_gibbs_supported(v0, W, bh=None, bv=None, u=None):
    """True if mnn_rbm_gibbs takes this chain: shape fits shared memory, operands 16-byte aligned, contiguous uniforms."""
    D, H = W.shape
    if not int(lib.mnn_rbm_gibbs_smem_bytes(D, H)) or not W.is_contiguous() or v0.dim() != 2 or v0.stride(1) != 1:
        return False
    if u is not None and not (torch.is_tensor(u[0]) and torch.is_tensor(u[1]) and u[0].dim() == 3 and u[1].dim() == 3):
        return False
    for t in (W, bh, bv) + (tuple(u) if u is not None else ()):
        if t is not None and (t.data_ptr() % 16 or (t.dim() == 2 and t.shape[0] > 1 and t.stride(0) % 4)
                              or t.stride(-1) != 1):
            return False
    if u is not None and not (u[0].is_contiguous() and u[1].is_contiguous()):
        return False
    return True


def rbm_gibbs(v0, W, bh, bv, k, p_v=None, v_k=None, h_k=None, u=None, seed=0, offset=0):
    """k-step Gibbs chain from v0[N,D] in one launch: fills p_v[N,D] (last step's probabilities), v_k[N,D], h_k[N,H].
    bh/bv: [N,.] per-row or [1,.]/[.] broadcast biases. u = (uh[k,N,H], uv[k,N,D]) or None (in-kernel Philox)."""
    N, D = v0.shape
    H = W.shape[1]
    uh, uv = (None, None) if u is None else (u[0], u[1])
    if u is not None:
        assert tuple(uh.shape[-3:]) == (k, N, H) and tuple(uv.shape[-3:]) == (k, N, D), "u = (uh[k,N,H], uv[k,N,D])"
    check(lib.mnn_rbm_gibbs(_ptr(v0), _rowstride(v0), _ptr(W), _ptr(bh), _bias_ld(bh, N), _ptr(bv), _bias_ld(bv, N),
                            _ptr(uh), _ptr(uv), int(u is None), seed, offset, _ptr(p_v),
                            _rowstride(p_v) if p_v is not None else 0, _ptr(v_k),
                            _rowstride(v_k) if v_k is not None else 0, _ptr(h_k),
                            _rowstride(h_k) if h_k is not None else 0, N, D, H, int(k), _stream()), "rbm_gibbs")


def sigmoid_bwd(y, dy, dpre):
    N, Cc = y.shape
    check(lib.mnn_sigmoid_bwd(_ptr(y), _rowstride(y), _ptr(dy), _rowstride(dy), _ptr(dpre), _rowstride(dpre), N, Cc,
                              _stream()), "sigmoid_bwd")


def rbm_free_energy(pre, bh, v, bv, F):
    N, H = pre.shape
    D = v.shape[1]
    ld_bh = bh.stride(0) if (bh.dim() == 2 and bh.shape[0] == N and N > 1) else 0
    ld_bv = bv.stride(0) if (bv.dim() == 2 and bv.shape[0] == N and N > 1) else 0
    check(lib.mnn_rbm_free_energy(_ptr(pre), _rowstride(pre), _ptr(bh), ld_bh, _ptr(v), _rowstride(v), _ptr(bv),
                                  ld_bv, _ptr(F), N, H, D, _stream()), "rbm_free_energy")


_ws = {}


def _reduce_ws(device):
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    if key not in _ws:
        _ws[key] = torch.empty(int(lib.mnn_reduce_workspace_bytes()), dtype=torch.uint8, device=device)
    return _ws[key]


def sum_into(x, out, scale=1.0, accumulate=False):
    check(lib.mnn_sum(_ptr(x), x.numel(), _ptr(_reduce_ws(x.device)), _ptr(out), float(scale), int(accumulate),
                      _stream()), "sum")


def sqnorm_into(x, out):
    check(lib.mnn_sqnorm(_ptr(x), x.numel(), _ptr(_reduce_ws(x.device)), _ptr(out), _stream()), "sqnorm")


def clip_adam(p, g, m, v, sqnorm, step, lr, grad_scale=1.0, clip_norm=5.0, beta1=0.9, beta2=0.999, eps=1e-4):
    check(lib.mnn_clip_adam(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(sqnorm), float(grad_scale),
                            float(clip_norm), float(lr), float(beta1), float(beta2), float(eps), int(step), _stream()),
          "clip_adam")


def scale_rows(x, w, ncols=None, period=None):
    """x[r, :ncols] *= w[r % period] on a row-major 2-D view x (row stride free)."""
    rows = x.shape[0]
    check(lib.mnn_scale_rows(_ptr(x), _rowstride(x), int(ncols if ncols is not None else x.shape[1]), _ptr(w), rows,
                             int(period if period is not None else w.numel()), _stream()), "scale_rows")


def axpy(y, x, alpha):
    """y += alpha * x (contiguous tensors of equal size)."""
    assert y.is_contiguous() and x.is_contiguous() and y.numel() == x.numel()
    check(lib.mnn_axpy(_ptr(y), _ptr(x), float(alpha), y.numel(), _stream()), "axpy")


def clip_sgd(p, g, sqnorm, lr, grad_scale=1.0, clip_norm=5.0):
    check(lib.mnn_clip_sgd(_ptr(p), _ptr(g), p.numel(), _ptr(sqnorm), float(grad_scale), float(clip_norm), float(lr),
                           _stream()), "clip_sgd")


def probe_mufu(kind='ex2', blocks=None, threads=512, iters=4096):
    """XU-pipe peak probe: returns MUFU instructions per second (per thread-lane ops/s) measured with CUDA events."""
    k = {'ex2': 0, 'rcp': 1, 'sigmoid': 2}[kind]
    blocks = blocks or 8 * num_sms()
    out = torch.zeros(blocks, device='cuda')
    check(lib.mnn_probe_mufu(_ptr(out), blocks, threads, 64, k, _stream()), "probe_mufu")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib.mnn_probe_mufu(_ptr(out), blocks, threads, iters, k, _stream()), "probe_mufu")
    e1.record()
    torch.cuda.synchronize()
    n = blocks * threads * iters * 8 * (2 if k == 2 else 1)
    return n / (e0.elapsed_time(e1) * 1e-3)


# ----------------------------------------------------------------------------- fused generation (K6)
# 'auto': the one-launch persistent kernel (csrc/gen_fused.cu) for Philox sampling when the shape fits (B <= 128, the
# sampled frame is the layer-0 input), the fp32-accurate multi-launch loop when uniforms are supplied (parity runs) or
# the shape does not fit; 'fused' / 'steps' force one or the other.
GENERATE_MODE = 'auto'
_gen_ws = {}


def generate_fused_supported(num_layers, num_inputs, units, B, M, D, H):
    r = list(units) + [0]
    return num_layers in (1, 2) and num_inputs == M * D and int(
        lib.mnn_generate_fused_workspace_bytes(num_layers, num_inputs, r[0], r[1], B, M, D, H)) > 0


def generate_fused(kernels, biases, state, dense_kernel, dense_bias, w_enc, w_dec, fc, out, u=None, use_philox=True, seed=0,
                   offset0=0):
    """All S = out.shape[1] generation steps in one launch. state: [(c, h)] per layer (updated in place), fc[B, C] the
    Dense output after the intro (updated to the one after the last step), out[B, S, M*D]."""
    L = len(kernels)
    B, S = out.shape[0], out.shape[1]
    M, D, H = w_enc.shape
    units = [k.shape[1] // 4 for k in kernels]
    num_inputs = kernels[0].shape[0] - units[0]
    r = units + [0]
    need = int(lib.mnn_generate_fused_workspace_bytes(L, num_inputs, r[0], r[1], B, M, D, H))
    if not need:
        raise ValueError('generate_fused: shape not taken (B <= 128, 1-2 layers, num_inputs == M*D)')
    key = (fc.device.index, torch.cuda.current_stream().cuda_stream, need)
    ws = _gen_ws.get(key)
    if ws is None:
        ws = _gen_ws[key] = torch.empty(need, dtype=torch.uint8, device=fc.device)
    assert out.is_contiguous() and fc.stride(1) == 1 and all(k.is_contiguous() for k in kernels)
    c1, h1 = (state[1] if L > 1 else (None, None))
    check(lib.mnn_generate_fused(L, num_inputs, _ptr(kernels[0]), _ptr(biases[0]), _ptr(state[0][0]), _ptr(state[0][1]),
                                 units[0], _ptr(kernels[1]) if L > 1 else None, _ptr(biases[1]) if L > 1 else None,
                                 _ptr(c1), _ptr(h1), r[1], _ptr(dense_kernel), _ptr(dense_bias), _ptr(w_enc), _ptr(w_dec),
                                 _ptr(fc), fc.stride(0), _ptr(u), int(use_philox and u is None), seed, offset0, _ptr(out),
                                 out.stride(0), out.stride(1), B, S, M, D, H, _ptr(ws), _stream()), "generate_fused")
