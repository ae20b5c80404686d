"""GPU: the tcgen05 3xTF32 GEMM against fp64 matmul for every operand-major combination, ragged shapes,
split-K, bias/beta epilogues and the a_exact (binary A) 2-product mode. fp32-level accuracy is the contract."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K, ta, tb, bias=False, beta=0.0, a_exact=False, seed=0, alpha=1.0):
    from multinn_b200 import ops
    rng = np.random.default_rng(seed + M + 3 * N + 7 * K)
    if a_exact:
        A = (rng.random((K, M) if ta else (M, K)) < 0.2).astype(np.float32)
    else:
        A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bv = rng.standard_normal(N).astype(np.float32) if bias else None
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = alpha * ((A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64))
    if bias:
        ref = ref + bv
    ref = ref + beta * C0
    C = torch.from_numpy(C0.copy()).cuda()
    ops.gemm(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), C, transA=bool(ta), transB=bool(tb),
             bias=None if bv is None else torch.from_numpy(bv).cuda(), alpha=alpha, beta=beta, a_exact=a_exact, mode='tc')
    torch.cuda.synchronize()
    got = C.cpu().numpy().astype(np.float64)
    scale = np.sqrt(K) if not a_exact else np.sqrt(0.2 * K)
    # the tensor core truncates its fp32 accumulator once per instruction: the error of the hi.hi chain grows with the
    # number of K=8 steps of one split (measured ~6e-8 per step); everything else is ~2^-21
    return float(np.abs(got - ref).max() / scale) / (1.0 + min(K, 4096) / 64.0)


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 256, 64), (256, 64, 96), (300, 200, 100), (77, 340, 420),
                                   (1000, 1700, 256), (2048, 2048, 512), (4224, 1024, 512), (700, 520, 932)])
def test_gemm_tc_majors(M, N, K, ta, tb):
    err = _run(M, N, K, ta, tb)
    assert err < 4e-6, err


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0)])
def test_gemm_tc_epilogues(ta, tb):
    assert _run(200, 340, 256, ta, tb, bias=True) < 4e-6
    assert _run(200, 340, 256, ta, tb, beta=1.0) < 4e-6
    assert _run(130, 84, 64, ta, tb, bias=True, beta=1.0, alpha=0.5) < 4e-6


def test_gemm_tc_split_k_weight_grad_shapes():
    # dW = X^T dG with the time*batch rows as the reduction dimension (both operands MN-major)
    assert _run(420, 2048, 16384, 1, 0) < 4e-6
    assert _run(256, 1700, 8192, 1, 0, beta=1.0) < 4e-6
    assert _run(512, 1024, 4096, 1, 0, bias=True) < 4e-6


def test_gemm_tc_binary_a_two_products():
    assert _run(512, 2048, 420, 0, 0, a_exact=True, bias=True) < 4e-6
    assert _run(420, 512, 4096, 1, 0, a_exact=True) < 4e-6


def test_gemm_tc_matches_f32_kernel_on_views():
    from multinn_b200 import ops
    rng = np.random.default_rng(1)
    big = torch.from_numpy(rng.standard_normal((300, 932)).astype(np.float32)).cuda()
    W = torch.from_numpy(rng.standard_normal((932, 2048)).astype(np.float32)).cuda()
    C1 = torch.empty(300, 2048, device='cuda')
    C2 = torch.empty(300, 2048, device='cuda')
    ops.gemm(big[:, :420], W[:420], C1, mode='tc')
    ops.gemm(big[:, :420], W[:420], C2, mode='f32')
    assert float((C1 - C2).abs().max()) < 3e-4
    ops.gemm(big[:, 420:], W[420:], C1, beta=1.0, mode='tc')
    ops.gemm(big[:, 420:], W[420:], C2, beta=1.0, mode='f32')
    assert float((C1 - C2).abs().max()) < 6e-4


def _run_pair(M, N, K, ta, tb, **kw):
    """The same check under the bf16-pair operand split (mnn_set_gemm_split(1), what the training step selects). Returns
    max |error| / sqrt(K) divided by its allowance: 6e-5 for the split itself (~2^-17 per product: the worst element of a
    K-term dot of unit normals sits near 4 sigma * 2^-17, a factor 2 of slack) plus _run's 4e-6 per 64 K-steps for the
    accumulator's per-instruction truncation."""
    from multinn_b200 import ops
    with ops.gemm_split('pair'):
        e = _run(M, N, K, ta, tb, **kw)
    steps = 1.0 + min(K, 4096) / 64.0
    return e * steps / (6e-5 + 4e-6 * steps)


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (300, 200, 100), (1000, 1700, 256), (2048, 2048, 512), (700, 520, 932)])
def test_gemm_tc_pair_split(M, N, K, ta, tb):
    """bf16 pairs x = x1 + x2 (16 mantissa bits): A1.B1 + A1.B2 + A2.B1 into ONE double-buffered TMEM accumulator."""
    assert _run_pair(M, N, K, ta, tb) < 1.0


def test_gemm_tc_pair_split_epilogues_and_split_k():
    assert _run_pair(512, 340, 256, 0, 0, bias=True) < 1.0
    assert _run_pair(512, 340, 256, 0, 1, beta=1.0) < 1.0
    assert _run_pair(420, 2048, 16384, 1, 0) < 1.0
    assert _run_pair(256, 1700, 8192, 1, 0, beta=1.0) < 1.0
    assert _run_pair(512, 2048, 420, 0, 0, a_exact=True, bias=True) < 1.0
    # many tiles per cluster: both accumulator buffers are reused several times
    assert _run_pair(8192, 2048, 96, 0, 0, bias=True) < 1.0


@pytest.mark.parametrize("tb", [0, 1])
@pytest.mark.parametrize("M,N,K,a_exact", [(512, 2048, 420, True), (1024, 1024, 512, False), (768, 1700, 256, False),
                                           (512, 256, 1700, False), (2048, 512, 1024, False), (300, 200, 100, False)])
def test_gemm_tc_presplit_weights(M, N, K, a_exact, tb):
    """B handed over as bf16 pair planes (mnn_split_bf16_pair + mnn_gemm_tc_bpair; TMA writes the operand tiles directly,
    K-major and MN-major): same result as the in-kernel split of the same pair mode (bit for bit without split-K) and within the pair bound
    of the fp64 product. Ragged N (1700: plane stride padded to 1704) and K (420, 100) included."""
    from multinn_b200 import ops
    rng = np.random.default_rng(M + N + K)
    A = ((rng.random((M, K)) < 0.2) if a_exact else rng.standard_normal((M, K))).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    Ad, Bd, bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(bias).cuda()
    C1, C2 = torch.empty(M, N, device='cuda'), torch.empty(M, N, device='cuda')
    with ops.gemm_split('pair'):
        ops.gemm(Ad, Bd, C1, transB=bool(tb), bias=bd, a_exact=a_exact)
        ops.gemm(Ad, Bd, C2, transB=bool(tb), bias=bd, a_exact=a_exact, b_weight=True)
        used = len(ops._pair_cache)
    assert used == (1 if (M >= 256 and N > 128 and K >= 64) else 0)
    torch.cuda.synchronize()
    if K < 1024:
        assert torch.equal(C1, C2)
    else:       # split-K: partial sums are red.add-ed in arrival order, two runs differ in the last bit
        assert float((C1 - C2).abs().max()) < 2e-4
    ref = A.astype(np.float64) @ (B.T if tb else B).astype(np.float64) + bias
    scale = np.sqrt(0.2 * K) if a_exact else np.sqrt(K)
    assert float(np.abs(C2.cpu().numpy() - ref).max() / scale) < 6e-5 + 4e-6 * (1 + K / 64)


@pytest.mark.parametrize("use_bpair", [False, True])
def test_gemm_tc_bf16_input_plane(use_bpair):
    """Binary stacked piano-roll rows as an exact bf16 plane (mnn_pack_stacked_bf16) feeding mnn_gemm_tc_abf16 as the A
    operand, K-major (layer-0 projection X.Wx over a row block) and MN-major (x-rows weight gradient X^T.dG): identical
    to the pair split of the fp32 rows, which is exact for binary A."""
    from multinn_b200 import ops
    rng = np.random.default_rng(5)
    Bb, T, D, Mt, N = 16, 63, 84, 5, 2048
    I = D * Mt
    x = (rng.random((Bb, T, D, Mt)) < 0.07).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    xin = torch.zeros((T + 1) * Bb, I, device='cuda')
    ops.pack_pianoroll(xd, xin.view(T + 1, Bb, I))
    twin = torch.empty((T + 1) * Bb, 424, dtype=torch.int16, device='cuda')
    ops.pack_stacked_bf16(xd.to(torch.uint8), twin)
    # the plane holds exactly the staged rows
    back = (twin.view(torch.bfloat16).float())[:, :I]
    assert torch.equal(back, xin) and float(twin[:, I:].abs().max()) == 0
    W = torch.from_numpy(rng.standard_normal((I, N)).astype(np.float32)).cuda()
    dG = torch.from_numpy(rng.standard_normal(((T + 1) * Bb, N)).astype(np.float32)).cuda()
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).cuda()
    rows = slice(Bb, Bb + 768)                         # a row block that does not start at the base
    C1, C2 = torch.empty(768, N, device='cuda'), torch.empty(768, N, device='cuda')
    G1, G2 = torch.empty(I, N, device='cuda'), torch.empty(I, N, device='cuda')
    with ops.gemm_split('pair'):
        ops.gemm(xin[rows], W, C1, bias=bias, a_exact=True, b_weight=use_bpair)
        ops.gemm(xin, dG, G1, transA=True, a_exact=True)
        n0 = len(ops._twins)
        ops.register_twin(xin, twin)
        ops.gemm(xin[rows], W, C2, bias=bias, a_exact=True, b_weight=use_bpair)
        ops.gemm(xin, dG, G2, transA=True, a_exact=True)
        assert n0 == 0 and len(ops._twins) == 1
    torch.cuda.synchronize()
    assert torch.equal(C1, C2)
    assert float((G1 - G2).abs().max()) < 2e-4 * float(G1.abs().max())          # split-K: arrival order of the partial sums
    ref = x.reshape(Bb, T, I).transpose(1, 0, 2).reshape(T * Bb, I)[0:768].astype(np.float64) @ W.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(C2.cpu().numpy(), ref + bias.cpu().numpy(), atol=6e-5 * np.sqrt(0.07 * I) * 3)
