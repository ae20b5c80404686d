"""CPU: the segment-row / four-GEMM formulation of NADE (tools/nade_segment_gemm.py, the design of round 2's tensor-core
NADE kernels) equals the reference's loop form (oracle, common/nade.py:155-229) in value and in every gradient."""
import importlib.util
import os

import numpy as np
import torch

from oracle import np_oracle as O
from oracle import torch_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('nade_segment_gemm', os.path.join(ROOT, 'tools', 'nade_segment_gemm.py'))
SG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(SG)


def _case(seed, N, D, H, density):
    rng = np.random.default_rng(seed)
    V = (rng.random((N, D)) < density).astype(np.float64)
    V[0] = 0.0                       # no set bit: one segment
    V[1] = 1.0                       # every bit set: D segments (the dense worst case)
    V[2] = 0.0
    V[2, D - 1] = 1.0                # only the last bit: it opens no segment
    V[3] = 0.0
    V[3, 0] = 1.0                    # first bit: segment 0 owns a single dim
    return (V, rng.standard_normal((N, H)) * 0.5, rng.standard_normal((N, D)) * 0.5,
            rng.standard_normal((D, H)) / np.sqrt(D), rng.standard_normal((D, H)) / np.sqrt(D))


def test_segment_rows_partition_every_dim_once():
    V = _case(0, 9, 20, 8, 0.2)[0]
    seg = SG.segment_rows(V)
    cover = np.zeros_like(V)
    for n, lo, hi in zip(seg['n'], seg['lo'], seg['hi']):
        cover[n, lo:hi + 1] += 1
    assert np.all(cover == 1)
    assert seg['n'].size == int(V[:, :-1].sum()) + V.shape[0]          # K_n + 1 segments per row
    assert np.all(V[seg['n'][seg['k'] > 0], seg['opener'][seg['k'] > 0]] == 1)


def test_four_gemm_form_equals_loop_form_and_autograd():
    for seed, (N, D, H, density) in enumerate([(12, 84, 32, 0.05), (8, 20, 16, 0.3), (6, 84, 64, 0.6)]):
        V, b_enc, b_dec, w_enc, w_dec = _case(seed, N, D, H, density)
        nll, P, cache = SG.forward(V, b_enc, b_dec, w_enc, w_dec)
        ref_nll, ref_P = O.nade_log_prob(V, b_enc, b_dec, w_enc, w_dec)
        np.testing.assert_allclose(nll, ref_nll, rtol=1e-12)
        np.testing.assert_allclose(P, ref_P, rtol=1e-12)
        dnll = np.random.default_rng(seed + 100).random(N)
        got = SG.backward(V, w_enc, w_dec, cache, dnll)
        t = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
        tb, td, te, tw = t(b_enc), t(b_dec), t(w_enc), t(w_dec)
        tn, _ = R.nade_log_prob(torch.tensor(V), tb, td, te, tw)
        (tn * torch.tensor(dnll)).sum().backward()
        for name, ref in (('b_enc', tb), ('b_dec', td), ('w_enc', te), ('w_dec', tw)):
            np.testing.assert_allclose(got[name], ref.grad.numpy(), rtol=1e-9, atol=1e-12, err_msg=name)


def test_flop_model():
    f = SG.flops(524288, 84, 256, 0.05 * 83, tracks=5)
    assert 1.3e7 < f['segment_rows'] < 1.4e7
    assert f['gemm_flops'] < 0.2 * f['dense_triangular_flops']


def test_tile_packing_keeps_rows_whole():
    rng = np.random.default_rng(5)
    V = (rng.random((3000, 84)) < 0.05).astype(np.float64)
    V[7] = 1.0                                            # 84 segments in one row
    starts, fill = SG.pack_tiles(V, 128)
    per_row = V[:, :-1].sum(1).astype(int) + 1
    for a, b in zip(starts[:-1], starts[1:]):
        assert 0 < per_row[a:b].sum() <= 128
    assert starts[0] == 0 and starts[-1] == 3000 and fill > 0.9


def test_tile_by_tile_schedule_equals_the_whole():
    """The planned kernels' order of work (tiles of whole source rows, the producer's k-block slices, per-lane epilogue
    ranges, in-tile suffix sums, persistent weight-gradient accumulators flushed every few tiles) gives the same values
    and gradients as the untiled four-GEMM form."""
    V, b_enc, b_dec, w_enc, w_dec = _case(7, 40, 84, 64, 0.08)
    dnll = np.random.default_rng(3).random(40)
    nll, P, cache = SG.forward(V, b_enc, b_dec, w_enc, w_dec)
    ref = SG.backward(V, w_enc, w_dec, cache, dnll)
    for tile_rows, kblock, flush in ((128, 32, 2), (96, 16, 1), (200, 64, 100)):
        n2, P2, g2 = SG.forward_backward_tiled(V, b_enc, b_dec, w_enc, w_dec, dnll, tile_rows, kblock, flush)
        np.testing.assert_allclose(n2, nll, rtol=1e-12)
        np.testing.assert_allclose(P2, P, rtol=1e-12)
        for name in ref:
            np.testing.assert_allclose(g2[name], ref[name], rtol=1e-10, atol=1e-13, err_msg=f'{name} {tile_rows}')
