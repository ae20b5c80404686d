"""GPU parity tests, kernel level: every C-ABI entry point against the CPU oracle on seeded inputs.
Tolerances: NLL / probabilities 1e-4 relative (BASELINE north_star, fp32 path); samples bit-exact."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import torch_ref as R

pytestmark = pytest.mark.gpu


def _ops():
    from multinn_b200 import ops
    return ops


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).cuda()


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ----------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,ta,tb", [(37, 53, 19, 0, 0), (256, 1700, 256, 0, 0), (130, 129, 70, 1, 0),
                                         (64, 420, 1024, 0, 1), (420, 2048, 8192, 1, 0), (1, 340, 256, 0, 0),
                                         (300, 84, 168, 1, 1)])
def test_gemm_f32(M, N, K, ta, tb):
    ops = _ops()
    rng = np.random.default_rng(M * 7 + N)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
    C = dev(C0)
    ops.gemm(dev(A), dev(B), C, transA=bool(ta), transB=bool(tb), mode='f32')
    assert rel_err(C.cpu().numpy(), ref) < 2e-6
    C = dev(C0)
    ops.gemm(dev(A), dev(B), C, transA=bool(ta), transB=bool(tb), bias=dev(bias), alpha=0.5, beta=1.0, mode='f32')
    assert rel_err(C.cpu().numpy(), 0.5 * ref + C0 + bias) < 2e-6


def test_gemm_strided_views():
    ops = _ops()
    rng = np.random.default_rng(3)
    big = dev(rng.standard_normal((100, 300)).astype(np.float32))
    A = big[:, 10:80]                      # row stride 300
    W = dev(rng.standard_normal((200, 64)).astype(np.float32))
    C = torch.zeros(100, 128, device='cuda')
    ops.gemm(A, W[50:120], C[:, 32:96], mode='f32')
    ref = A.double().cpu().numpy() @ W[50:120].double().cpu().numpy()
    assert rel_err(C[:, 32:96].cpu().numpy(), ref) < 2e-6
    assert float(C[:, :32].abs().max()) == 0 and float(C[:, 96:].abs().max()) == 0


# ----------------------------------------------------------------------------- input staging
def test_pack_pianoroll():
    ops = _ops()
    B, T, D, M = 5, 7, 84, 5
    x = O.synthetic_pianoroll(B, T, D, M, density=0.2, seed=1)
    xin = torch.full((T + 1, B, D * M), -1.0, device='cuda')
    xtr = torch.full((M, T + 1, B, D), -1.0, device='cuda')
    bits = torch.zeros(M, T * B, 4, dtype=torch.int32, device='cuda')
    ops.pack_pianoroll(dev(x), xin, xtr, bits)
    inp, tgt = O.composer_inputs_targets(x)             # [B,T,D*M]
    np.testing.assert_array_equal(xin[:T].cpu().numpy(), inp.transpose(1, 0, 2))
    np.testing.assert_array_equal(xin[1:].cpu().numpy(), tgt.transpose(1, 0, 2))
    for m in range(M):
        np.testing.assert_array_equal(xtr[m, 1:].cpu().numpy(), x[..., m].transpose(1, 0, 2))
        assert float(xtr[m, 0].abs().max()) == 0
    b = bits.cpu().numpy().view(np.uint32)
    for m in range(M):
        for t in range(T):
            for bb in range(B):
                row = x[bb, t, :, m]
                words = [sum(int(row[w * 32 + j]) << j for j in range(32) if w * 32 + j < D) for w in range(4)]
                assert list(b[m, t * B + bb]) == words


# ----------------------------------------------------------------------------- NADE
def _nade_case(N, D, H, M, seed, density=0.1):
    rng = np.random.default_rng(seed)
    x = (rng.random((M, N, D)) < density).astype(np.float32)
    U = M * (H + D)
    fc = (rng.standard_normal((N, U)) * 0.7).astype(np.float32)
    std = 1 / np.sqrt(D)
    we = (rng.standard_normal((M, D, H)) * std).astype(np.float32)
    wd = (rng.standard_normal((M, D, H)) * std).astype(np.float32)
    return x, fc, we, wd


def _bits_of(x_m):
    ops = _ops()
    M, N, D = x_m.shape
    bits = torch.empty(M, N, 4, dtype=torch.int32, device='cuda')
    for m in range(M):
        ops.pack_rows(dev(x_m[m]), bits[m], D)
    return bits


@pytest.fixture(params=['simt', 'tc'])
def nade_mode(request):
    """Both forward kernels behind mnn_nade_logprob_fwd: the SIMT segment kernel (nade.cu, default) and the tcgen05
    segment-row kernel (nade_tc.cu); same tolerances."""
    ops = _ops()
    ops.set_nade_mode(request.param)
    yield request.param
    ops.set_nade_mode('simt')


@pytest.mark.parametrize("N,D,H,M,density", [(301, 84, 256, 5, 0.05), (64, 84, 256, 1, 0.5), (130, 84, 128, 2, 1.0),
                                             (33, 20, 128, 3, 0.0), (1000, 84, 256, 5, 0.2), (4000, 84, 256, 5, 0.01)])
def test_nade_logprob_fwd(N, D, H, M, density, nade_mode):
    ops = _ops()
    x, fc, we, wd = _nade_case(N, D, H, M, seed=N + D, density=density)
    bits = _bits_of(x)
    nll = torch.empty(M, N, device='cuda')
    cp = torch.empty(M, N, D, device='cuda')
    ops.nade_logprob_fwd(bits, dev(fc), 0, M * H, dev(we), dev(wd), nll, cond_p=cp)
    f64 = np.float64
    for m in range(M):
        be, bd = fc[:, m * H:(m + 1) * H], fc[:, M * H + m * D:M * H + (m + 1) * D]
        ref_nll, ref_p = O.nade_log_prob(x[m].astype(f64), be.astype(f64), bd.astype(f64), we[m].astype(f64),
                                         wd[m].astype(f64))
        np.testing.assert_allclose(nll[m].cpu().numpy(), ref_nll, rtol=1e-4)
        np.testing.assert_allclose(cp[m].cpu().numpy(), ref_p, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("N,D,H,M,density", [(257, 84, 256, 5, 0.05), (50, 84, 128, 2, 0.6), (40, 20, 128, 3, 0.3)])
def test_nade_logprob_bwd(N, D, H, M, density, nade_mode):
    ops = _ops()
    x, fc, we, wd = _nade_case(N, D, H, M, seed=N * 3 + H, density=density)
    bits = _bits_of(x)
    gscale = 1.0 / (N * M)
    # oracle: torch autograd through the unrolled loop, fp64
    t = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    fc_t, we_t, wd_t = t(fc), t(we), t(wd)
    loss = 0
    for m in range(M):
        nll_m, _ = R.nade_log_prob(torch.tensor(x[m], dtype=torch.float64), fc_t[:, m * H:(m + 1) * H],
                                   fc_t[:, M * H + m * D:M * H + (m + 1) * D], we_t[m], wd_t[m])
        loss = loss + nll_m.sum() * gscale
    loss.backward()
    fcd, wed, wdd = dev(fc), dev(we), dev(wd)
    nll = torch.empty(M, N, device='cuda')
    dfc = torch.zeros_like(fcd)
    dwe, dwd = torch.zeros_like(wed), torch.zeros_like(wdd)
    ops.nade_logprob_fwd(bits, fcd, 0, M * H, wed, wdd, nll, dfc=dfc, gscale=gscale)
    ops.nade_logprob_bwd(bits, fcd, 0, M * H, wed, wdd, dfc, dwe, dwd)
    assert rel_err(dfc.cpu().numpy(), fc_t.grad.numpy()) < 1e-4
    assert rel_err(dwe.cpu().numpy(), we_t.grad.numpy()) < 1e-4
    assert rel_err(dwd.cpu().numpy(), wd_t.grad.numpy()) < 1e-4


@pytest.mark.parametrize("budget", [0, 20])
def test_nade_row_chunks_and_sm_budget_equal_whole_call(budget, nade_mode):
    """The chunk pipeline calls the NADE kernels on row slices of the [M,N,..] buffers (track_stride > rows) under an SM
    budget: forward results are bit-identical to one whole call, accumulated weight gradients equal within fp32 order."""
    ops = _ops()
    N, D, H, M = 300, 84, 256, 5
    x, fc, we, wd = _nade_case(N, D, H, M, seed=77, density=0.07)
    bits = _bits_of(x)
    fcd, wed, wdd = dev(fc), dev(we), dev(wd)
    gscale = 1.0 / (N * M)

    def run(chunks):
        nll = torch.zeros(M, N, device='cuda')
        dfc = torch.zeros_like(fcd)
        dwe, dwd = torch.zeros_like(wed), torch.zeros_like(wdd)
        ops.set_sm_budget(budget)
        try:
            for r0, r1 in chunks:
                ops.nade_logprob_fwd(bits[:, r0:r1], fcd[r0:r1], 0, M * H, wed, wdd, nll[:, r0:r1], dfc=dfc[r0:r1], gscale=gscale)
                ops.nade_logprob_bwd(bits[:, r0:r1], fcd[r0:r1], 0, M * H, wed, wdd, dfc[r0:r1], dwe, dwd)
        finally:
            ops.set_sm_budget(0)
        torch.cuda.synchronize()
        return nll, dfc, dwe, dwd
    whole = run([(0, N)])
    parts = run([(0, 96), (96, 97), (97, 300)])
    assert torch.equal(whole[0], parts[0]) and torch.equal(whole[1], parts[1])
    assert rel_err(parts[2].cpu().numpy(), whole[2].cpu().numpy()) < 1e-6
    assert rel_err(parts[3].cpu().numpy(), whole[3].cpu().numpy()) < 1e-6


@pytest.mark.parametrize("N,D,H,M", [(200, 84, 256, 5), (37, 84, 256, 1), (64, 20, 128, 2)])
def test_nade_sample_bit_exact(N, D, H, M):
    ops = _ops()
    _, fc, we, wd = _nade_case(N, D, H, M, seed=11 * N)
    rng = np.random.default_rng(5)
    u = rng.random((M, N, D), dtype=np.float32)
    out = torch.empty(N, D * M, device='cuda')
    nll = torch.empty(M, N, device='cuda')
    ops.nade_sample(dev(fc), 0, M * H, dev(we), dev(wd), out, D * M, M, 1, u=dev(u), nll=nll)
    got = out.cpu().numpy().reshape(N, D, M)
    f64 = np.float64
    for m in range(M):
        be, bd = fc[:, m * H:(m + 1) * H], fc[:, M * H + m * D:M * H + (m + 1) * D]
        v, ref_nll = O.nade_sample(be.astype(f64), bd.astype(f64), we[m].astype(f64), wd[m].astype(f64),
                                   u[m].astype(f64))
        np.testing.assert_array_equal(got[:, :, m], v)
        np.testing.assert_allclose(nll[m].cpu().numpy(), ref_nll, rtol=1e-4)
    # temperature=None -> threshold
    ops.nade_sample(dev(fc), 0, M * H, dev(we), dev(wd), out, D * M, M, 1, u=None, use_philox=False)
    got = out.cpu().numpy().reshape(N, D, M)
    for m in range(M):
        be, bd = fc[:, m * H:(m + 1) * H], fc[:, M * H + m * D:M * H + (m + 1) * D]
        v, _ = O.nade_sample(be.astype(f64), bd.astype(f64), we[m].astype(f64), wd[m].astype(f64), None)
        np.testing.assert_array_equal(got[:, :, m], v)


def test_nade_sample_philox_statistics():
    """Philox path: samples are {0,1}, deterministic for a (seed, offset), and consistent with the model:
    v == (u_implied < p) cannot be checked without u, so check mean(v) ~ mean(p) under zero weights."""
    ops = _ops()
    N, D, H, M = 4096, 84, 256, 2
    fc = np.zeros((N, M * (H + D)), np.float32)
    fc[:, M * H:] = -1.0                                  # p = sigmoid(-1 + 0.5*sum(w_dec)) with zero w -> sigmoid(-1)
    we = np.zeros((M, D, H), np.float32)
    wd = np.zeros((M, D, H), np.float32)
    out1 = torch.empty(N, D * M, device='cuda')
    out2 = torch.empty(N, D * M, device='cuda')
    ops.nade_sample(dev(fc), 0, M * H, dev(we), dev(wd), out1, D * M, M, 1, use_philox=True, seed=7, offset=3)
    ops.nade_sample(dev(fc), 0, M * H, dev(we), dev(wd), out2, D * M, M, 1, use_philox=True, seed=7, offset=3)
    assert torch.equal(out1, out2)
    ops.nade_sample(dev(fc), 0, M * H, dev(we), dev(wd), out2, D * M, M, 1, use_philox=True, seed=7, offset=4)
    assert not torch.equal(out1, out2)
    v = out1.cpu().numpy()
    assert set(np.unique(v)) <= {0.0, 1.0}
    p = 1 / (1 + np.exp(1.0))
    assert abs(v.mean() - p) < 4 * np.sqrt(p * (1 - p) / v.size)


# ----------------------------------------------------------------------------- LSTM
def _lstm_case(T, B, I, R, seed):
    rng = np.random.default_rng(seed)
    x = (rng.random((T, B, I)) < 0.1).astype(np.float32)
    kernel = O._glorot(rng, I + R, 4 * R)
    bias = (rng.standard_normal(4 * R) * 0.1).astype(np.float32)
    return x, kernel, bias


@pytest.mark.parametrize("mode,persistent", [("simt", False), ("tc", False), ("tc", True)])
@pytest.mark.parametrize("T,B,I,Rn,keep", [(9, 5, 84, 64, 1.0), (6, 33, 420, 512, 1.0), (7, 4, 30, 32, 0.8),
                                           (12, 300, 64, 256, 0.9), (5, 130, 20, 48, 1.0), (3, 1100, 16, 96, 1.0), (4, 1024, 24, 512, 0.9), (5, 256, 24, 256, 0.9),
                                           (3, 512, 16, 256, 1.0)])
def test_lstm_sequence_fwd_bwd(T, B, I, Rn, keep, mode, persistent):
    ops = _ops()
    x, kernel, bias = _lstm_case(T, B, I, Rn, seed=T * B)
    rng = np.random.default_rng(9)
    u = rng.random((T, B, Rn), dtype=np.float32)
    dout = rng.standard_normal((T, B, Rn)).astype(np.float32)
    # oracle (fp64 autograd)
    xt = torch.tensor(x.transpose(1, 0, 2), dtype=torch.float64)          # [B,T,I]
    k_t = torch.tensor(kernel, dtype=torch.float64, requires_grad=True)
    b_t = torch.tensor(bias, dtype=torch.float64, requires_grad=True)
    outs, state = R.rnn_scan(xt, [(k_t, b_t)], keep, [torch.tensor(u, dtype=torch.float64)] if keep < 1 else None)
    (outs * torch.tensor(dout.transpose(1, 0, 2), dtype=torch.float64)).sum().backward()
    # device
    xd, kd, bd = dev(x), dev(kernel), dev(bias)
    gates = torch.empty(T, B, 4 * Rn, device='cuda')
    ops.gemm(xd.view(T * B, I), kd[:I], gates.view(T * B, 4 * Rn), bias=bd, mode='f32')
    hbuf = torch.zeros(T + 1, B, Rn, device='cuda')
    cbuf = torch.zeros(T + 1, B, Rn, device='cuda')
    out = torch.empty(T, B, Rn, device='cuda')
    dscale = torch.empty(T, B, Rn, device='cuda')
    ops.lstm_seq_fwd(gates, kd[I:], hbuf, cbuf, out=out, dscale=dscale, u=dev(u), keep=keep, mode=mode,
                     persistent=persistent)
    assert rel_err(out.cpu().numpy(), outs.detach().numpy().transpose(1, 0, 2)) < 1e-5
    assert rel_err(cbuf[T].cpu().numpy(), state[0][0].detach().numpy()) < 1e-5
    dh_work = torch.empty(B, Rn, device='cuda')
    dc_work = torch.empty(B, Rn, device='cuda')
    ops.lstm_seq_bwd(gates, kd[I:], cbuf, dev(dout), dscale if keep < 1 else None, dh_work, dc_work, mode=mode,
                     persistent=persistent)
    dk = torch.empty_like(kd)
    db = torch.empty_like(bd)
    dg = gates.view(T * B, 4 * Rn)
    ops.gemm(xd.view(T * B, I), dg, dk[:I], transA=True, mode='f32')
    ops.gemm(hbuf[:T].view(T * B, Rn), dg, dk[I:], transA=True, mode='f32')
    ops.colsum(dg, db)
    assert rel_err(dk.cpu().numpy(), k_t.grad.numpy()) < 2e-5
    assert rel_err(db.cpu().numpy(), b_t.grad.numpy()) < 2e-5


@pytest.mark.parametrize("T,B,Rn", [(6, 8, 64), (5, 260, 512), (4, 512, 256)])
def test_lstm_dropout_philox_stream_is_global_row_keyed(T, B, Rn):
    """the dropout noise is a function of (global batch row, unit, global time step, seed) only. The SIMT
    cell, the 1-CTA and pair tcgen05 kernels draw the SAME masks, equal to the CPU Philox (oracle/philox.py::dropout_uniforms);
    a batch shard run under ops.row_map and a time chunk run with t_base reproduce the slices of the whole run."""
    from oracle.philox import dropout_uniforms
    ops = _ops()
    keep, seed = 0.8, 4242
    rng = np.random.default_rng(B)
    gates0 = dev(rng.standard_normal((T, B, 4 * Rn)).astype(np.float32))
    wh = dev((rng.standard_normal((Rn, 4 * Rn)) * 0.05).astype(np.float32))
    ref = np.floor(np.float32(keep) + dropout_uniforms(seed, T, B, Rn)) / np.float32(keep)

    def run(g, mode, persistent, t_base=0):
        Tn, Bn = g.shape[0], g.shape[1]
        g = g.clone()
        hbuf, cbuf = torch.zeros(Tn + 1, Bn, Rn, device='cuda'), torch.zeros(Tn + 1, Bn, Rn, device='cuda')
        out, dscale = torch.empty(Tn, Bn, Rn, device='cuda'), torch.empty(Tn, Bn, Rn, device='cuda')
        ops.lstm_seq_fwd(g, wh, hbuf, cbuf, out=out, dscale=dscale, keep=keep, seed=seed, mode=mode,
                         persistent=persistent, t_base=t_base)
        return dscale.cpu().numpy()

    for mode, persistent in (("simt", False), ("tc", False), ("tc", True)):
        np.testing.assert_array_equal(run(gates0, mode, persistent), ref, err_msg=f'{mode} {persistent}')
    half = B // 2
    with ops.row_map(half, B, half):                                   # second shard of a 2-way data-parallel split
        np.testing.assert_array_equal(run(gates0[:, half:].contiguous(), "tc", True), ref[:, half:])
    np.testing.assert_array_equal(run(gates0[2:].contiguous(), "tc", True, t_base=2), ref[2:])   # a later time chunk
    np.testing.assert_array_equal(run(gates0, "tc", True), ref)        # the map was restored


def test_half_step_and_gibbs_row_map_reproduce_global_run():
    """Philox Bernoulli draws of a row shard under ops.row_map equal the same rows of the unsharded call
    (time-major groups: local rows t*Bl + b <-> global rows t*Bg + base + b)."""
    ops = _ops()
    rng = np.random.default_rng(3)
    Tn, Bg, Bl, C = 3, 8, 4, 84
    pre = dev(rng.standard_normal((Tn * Bg, C)).astype(np.float32))
    s_all = torch.empty(Tn * Bg, C, device='cuda')
    ops.bias_sigmoid_sample(pre, s=s_all, use_philox=True, seed=5, offset=1000)
    for base in (0, 4):
        loc = pre.view(Tn, Bg, C)[:, base:base + Bl].reshape(Tn * Bl, C).contiguous()
        s_loc = torch.empty(Tn * Bl, C, device='cuda')
        with ops.row_map(Bl, Bg, base):
            ops.bias_sigmoid_sample(loc, s=s_loc, use_philox=True, seed=5, offset=1000)
        assert torch.equal(s_loc.view(Tn, Bl, C), s_all.view(Tn, Bg, C)[:, base:base + Bl])
    D, H, k = 84, 64, 2
    W = dev(O._glorot(rng, D, H))
    v0 = dev((rng.random((Tn * Bg, D)) < 0.2).astype(np.float32))
    bh, bv = torch.zeros(1, H, device='cuda'), torch.zeros(1, D, device='cuda')
    vk_all = torch.empty(Tn * Bg, D, device='cuda')
    ops.rbm_gibbs(v0, W, bh, bv, k, v_k=vk_all, seed=9, offset=77)
    loc = v0.view(Tn, Bg, D)[:, 4:].reshape(Tn * Bl, D).contiguous()
    vk_loc = torch.empty(Tn * Bl, D, device='cuda')
    with ops.row_map(Bl, Bg, 4):
        ops.rbm_gibbs(loc, W, bh, bv, k, v_k=vk_loc, seed=9, offset=77)
    assert torch.equal(vk_loc.view(Tn, Bl, D), vk_all.view(Tn, Bg, D)[:, 4:])


# ----------------------------------------------------------------------------- RBM half-steps, free energy
def test_bias_sigmoid_sample_and_free_energy():
    ops = _ops()
    rng = np.random.default_rng(4)
    N, D, H = 77, 84, 256
    v = (rng.random((N, D)) < 0.3).astype(np.float32)
    W = O._glorot(rng, D, H)
    bh = (rng.standard_normal((N, H)) * 0.3).astype(np.float32)
    bh1 = (rng.standard_normal((1, H)) * 0.3).astype(np.float32)
    bv1 = (rng.standard_normal((1, D)) * 0.3).astype(np.float32)
    u = rng.random((N, H), dtype=np.float32)
    pre = torch.empty(N, H, device='cuda')
    ops.gemm(dev(v), dev(W), pre)
    p = torch.empty(N, H, device='cuda')
    s = torch.empty(N, H, device='cuda')
    ops.bias_sigmoid_sample(pre, bias=dev(bh), u=dev(u), p=p, s=s)
    rp, rs = O.rbm_forward(v.astype(np.float64), W.astype(np.float64), bh.astype(np.float64), u.astype(np.float64))
    np.testing.assert_allclose(p.cpu().numpy(), rp, rtol=1e-5)
    np.testing.assert_array_equal(s.cpu().numpy(), rs)
    ops.bias_sigmoid_sample(pre, bias=dev(bh1), u=dev(u), p=p, s=s)
    rp, rs = O.rbm_forward(v.astype(np.float64), W.astype(np.float64), bh1.astype(np.float64), u.astype(np.float64))
    np.testing.assert_allclose(p.cpu().numpy(), rp, rtol=1e-5)
    np.testing.assert_array_equal(s.cpu().numpy(), rs)
    F = torch.empty(N, device='cuda')
    ops.rbm_free_energy(pre, dev(bh1), dev(v), dev(bv1), F)
    ref = O.rbm_free_energy(v.astype(np.float64), W.astype(np.float64), bh1.astype(np.float64), bv1.astype(np.float64))
    np.testing.assert_allclose(F.cpu().numpy(), ref, rtol=1e-5)


# ----------------------------------------------------------------------------- optimiser
@pytest.mark.parametrize("gnorm_scale", [0.01, 100.0])
def test_clip_adam_matches_tf_adam(gnorm_scale):
    ops = _ops()
    rng = np.random.default_rng(8)
    n = 100003
    p0 = rng.standard_normal(n).astype(np.float32)
    p, m, v = p0.astype(np.float64), np.zeros(n), np.zeros(n)
    pd, md, vd = dev(p0), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    sq = torch.zeros(1, device='cuda')
    for t in range(1, 4):
        g = (rng.standard_normal(n) * gnorm_scale).astype(np.float32)
        (gc,), gn = O.clip_by_global_norm([g.astype(np.float64)], 5.0)
        p, m, v = O.tf_adam_step(p, gc, m, v, t)
        gd = dev(g)
        ops.sqnorm_into(gd, sq)
        assert abs(float(sq.sqrt()) - gn) / gn < 1e-5
        ops.clip_adam(pd, gd, md, vd, sq, t, 0.01)
    assert rel_err(pd.cpu().numpy(), p) < 1e-5
    # SGD
    pd2 = dev(p0)
    ops.clip_sgd(pd2, gd, sq, 0.01)
    assert rel_err(pd2.cpu().numpy(), p0 - 0.01 * gc) < 1e-5
