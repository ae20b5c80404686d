"""Known-answer and cross-restatement tests that pin the oracle; CPU only."""
import itertools

import numpy as np
import torch

from oracle import np_oracle as O
from oracle import torch_ref as R


def _rand_nade(N=7, D=9, H=5, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    x = (rng.random((N, D)) < 0.3).astype(dtype)
    return (x, rng.standard_normal((N, H)).astype(dtype), rng.standard_normal((N, D)).astype(dtype),
            rng.standard_normal((D, H)).astype(dtype) * 0.5, rng.standard_normal((D, H)).astype(dtype) * 0.5)


def test_nade_loop_equals_triangular():
    args = _rand_nade()
    nll1, p1 = O.nade_log_prob(*args)
    nll2, p2 = O.nade_log_prob_triangular(*args)
    np.testing.assert_allclose(nll1, nll2, rtol=1e-13)
    np.testing.assert_allclose(p1, p2, rtol=1e-13)


def test_nade_zero_weights_closed_form():
    x, be, bd, we, wd = _rand_nade()
    nll, p = O.nade_log_prob(x, be, bd, we * 0, wd * 0)
    ps = 1 / (1 + np.exp(-bd))
    np.testing.assert_allclose(p, ps, rtol=1e-14)
    ref = -(x * np.log(1e-6 + ps) + (1 - x) * np.log(1e-6 + 1 - ps)).sum(1)
    np.testing.assert_allclose(nll, ref, rtol=1e-13)


def test_nade_normalises_up_to_safe_log_eps():
    D, H = 6, 4
    rng = np.random.default_rng(1)
    we, wd = rng.standard_normal((D, H)), rng.standard_normal((D, H))
    be, bd = rng.standard_normal((1, H)), rng.standard_normal((1, D))
    vs = np.array(list(itertools.product([0., 1.], repeat=D)))
    nll, _ = O.nade_log_prob(vs, np.repeat(be, len(vs), 0), np.repeat(bd, len(vs), 0), we, wd)
    total = np.exp(-nll).sum()
    assert abs(total - 1.0) < 5 * D * 1e-6 and total > 1.0


def test_nade_sample_consistent_with_log_prob():
    x, be, bd, we, wd = _rand_nade(N=16)
    u = np.random.default_rng(2).random(x.shape)
    v, nll_s = O.nade_sample(be, bd, we, wd, u)
    nll, p = O.nade_log_prob(v, be, bd, we, wd)
    np.testing.assert_allclose(nll, nll_s, rtol=1e-13)
    assert np.array_equal(v, (u < p).astype(v.dtype))
    v2, _ = O.nade_sample(be, bd, we, wd, u)
    assert np.array_equal(v, v2)                      # pure function of the uniforms
    vt, _ = O.nade_sample(be, bd, we, wd, None)        # temperature=None -> threshold
    _, pt = O.nade_log_prob(vt, be, bd, we, wd)
    assert np.array_equal(vt, (pt >= .5).astype(vt.dtype))


def test_torch_fp32_matches_numpy_fp64_nade():
    args = _rand_nade(N=32, D=84, H=64)
    nll64, p64 = O.nade_log_prob(*args)
    t = [torch.tensor(a, dtype=torch.float32) for a in args]
    nll32, p32 = R.nade_log_prob(*t)
    np.testing.assert_allclose(nll32.numpy(), nll64, rtol=1e-5)
    np.testing.assert_allclose(p32.numpy(), p64, atol=2e-6)


def test_lstm_cell_matches_torch_lstm_after_gate_permutation():
    rng = np.random.default_rng(3)
    B, I, Rr, T = 4, 6, 5, 7
    kernel = rng.standard_normal((I + Rr, 4 * Rr)) * 0.3
    bias = rng.standard_normal(4 * Rr) * 0.1
    xs = rng.standard_normal((B, T, I))
    outs, st = O.rnn_scan(xs, [(kernel, bias)])
    lstm = torch.nn.LSTM(I, Rr, batch_first=True).double()
    # TF order i,j,f,o -> torch order i,f,g(j),o
    perm = np.concatenate([np.arange(0, Rr), np.arange(2 * Rr, 3 * Rr), np.arange(Rr, 2 * Rr), np.arange(3 * Rr, 4 * Rr)])
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(torch.tensor(kernel[:I, perm].T))
        lstm.weight_hh_l0.copy_(torch.tensor(kernel[I:, perm].T))
        lstm.bias_ih_l0.copy_(torch.tensor(bias[perm]))
        lstm.bias_hh_l0.zero_()
        y, (hn, cn) = lstm(torch.tensor(xs))
    np.testing.assert_allclose(outs, y.numpy(), atol=1e-12)
    np.testing.assert_allclose(st[0][0], cn[0].numpy(), atol=1e-12)


def test_dropout_tf_semantics():
    x = np.ones((2, 4), np.float32)
    u = np.array([[0.0, 0.099, 0.1, 0.5], [0.9, 0.95, 0.1000001, 0.0999]], np.float32)
    y = O.dropout(x, 0.9, u)
    keep = np.floor(np.float32(0.9) + u)
    np.testing.assert_array_equal(y, x / np.float32(0.9) * keep)
    assert y[0, 0] == 0 and y[0, 3] > 1.1


def test_rbm_free_energy_vs_bruteforce():
    rng = np.random.default_rng(4)
    D, H = 5, 6
    W, bh, bv = rng.standard_normal((D, H)), rng.standard_normal((1, H)), rng.standard_normal((1, D))
    v = (rng.random((3, D)) < 0.5).astype(np.float64)
    hs = np.array(list(itertools.product([0., 1.], repeat=H)))
    E = -(v @ W @ hs.T) - (hs @ bh.T).T - (v @ bv.T)          # [3, 2^H]
    F_bf = -np.log(np.exp(-E).sum(1))
    np.testing.assert_allclose(O.rbm_free_energy(v, W, bh, bv), F_bf, rtol=1e-12)


def test_rbm_free_energy_cost_equals_broadcast_mean():
    rng = np.random.default_rng(5)
    D, H, N = 5, 6, 4
    W, bh, bv = rng.standard_normal((D, H)), rng.standard_normal((1, H)), rng.standard_normal((1, D))
    v = (rng.random((N, D)) < 0.5).astype(np.float64)
    vs = (rng.random((N, D)) < 0.5).astype(np.float64)

    def F_ref(vv):  # literal common/rbm.py:258 with its [N]-[N,1] -> [N,N] broadcast
        return -np.log(1 + np.exp(vv @ W + bh)).sum(1) - vv @ bv.T
    lit = (F_ref(v) - F_ref(vs)).mean()
    np.testing.assert_allclose(O.rbm_free_energy_cost_mean(v, vs, W, bh, bv), lit, rtol=1e-12)


def test_gibbs_pure_function_and_k0():
    rng = np.random.default_rng(6)
    D, H, N, k = 8, 7, 5, 3
    W, bh, bv = rng.standard_normal((D, H)), rng.standard_normal((N, H)), rng.standard_normal((N, D))
    v = (rng.random((N, D)) < 0.5).astype(np.float64)
    uh, uv = rng.random((k, N, H)), rng.random((k, N, D))
    a = O.rbm_gibbs(v, W, bh, bv, k, uh, uv)
    b = O.rbm_gibbs(v, W, bh, bv, k, uh, uv)
    assert np.array_equal(a[1], b[1]) and set(np.unique(a[1])) <= {0., 1.}
    p0, v0 = O.rbm_gibbs(v, W, bh, bv, 0, uh, uv)
    assert np.array_equal(p0, v) and np.array_equal(v0, v)


def test_tf_adam_first_step_closed_form_and_clip():
    p, g = np.array([1.0, -2.0]), np.array([0.5, -0.25])
    p1, m1, v1 = O.tf_adam_step(p, g, np.zeros(2), np.zeros(2), 1)
    lr_t = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
    np.testing.assert_allclose(p1, p - lr_t * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-4), rtol=1e-14)
    for scale, expect in ((1.0, 1.0), (10.0, None)):
        gs = [np.array([3.0, 4.0]) * scale]
        c, gn = O.clip_by_global_norm(gs, 5.0)
        assert np.isclose(gn, 5 * scale)
        assert np.isclose(np.linalg.norm(c[0]), 5.0)
    c, gn = O.clip_by_global_norm([np.array([0.3, 0.4])], 5.0)
    np.testing.assert_allclose(c[0], [0.3, 0.4])


def test_composer_torch_matches_numpy_and_fd_gradient():
    B, T, D, M, H = 2, 3, 5, 2, 4
    x = O.synthetic_pianoroll(B, T, D, M, density=0.3)
    p = O.init_composer_params(D, M, H, (6, 4), seed=7)
    out = O.composer_forward(x.astype(np.float64), O.cast_params(p, np.float64))
    tp = R.to_torch(p, torch.float64, requires_grad=True)
    loss, nll = R.composer_loss(torch.tensor(x, dtype=torch.float64), tp)
    np.testing.assert_allclose(float(loss), out['loss'], rtol=1e-12)
    np.testing.assert_allclose(nll.detach().numpy(), out['nll'], rtol=1e-12)
    leaves = R.flat_params(tp)
    grads = torch.autograd.grad(loss, leaves)
    # finite differences on a few entries of each leaf
    rng = np.random.default_rng(0)
    for leaf, g in zip(leaves, grads):
        flat = leaf.detach().view(-1)
        for idx in rng.choice(flat.numel(), size=min(3, flat.numel()), replace=False):
            old = float(flat[idx])
            with torch.no_grad():
                flat[idx] = old + 1e-6
                lp = float(R.composer_loss(torch.tensor(x, dtype=torch.float64), tp)[0])
                flat[idx] = old - 1e-6
                lm = float(R.composer_loss(torch.tensor(x, dtype=torch.float64), tp)[0])
                flat[idx] = old
            fd = (lp - lm) / 2e-6
            assert abs(fd - float(g.view(-1)[idx])) <= 1e-6 + 1e-5 * abs(fd)


def test_flatten_row_order_is_batch_major():
    x = O.synthetic_pianoroll(3, 4, 2, 2, density=0.5)
    inp, tgt = O.composer_inputs_targets(x)
    assert inp.shape == (3, 4, 4) and np.all(inp[:, 0] == 0)
    np.testing.assert_array_equal(inp[:, 1:], tgt[:, :-1])
    # feature index = d*M + m (multinn_composer.py:74-78)
    assert tgt[1, 2, 1 * 2 + 0] == x[1, 2, 1, 0]
    assert tgt.reshape(12, 4)[1 * 4 + 2, 3] == x[1, 2, 1, 1]


def test_composer_generate_shapes_and_determinism():
    B, Ti, D, M, H, S = 2, 3, 5, 2, 4, 4
    x = O.synthetic_pianoroll(B, Ti, D, M, density=0.3).astype(np.float64)
    p = O.cast_params(O.init_composer_params(D, M, H, (6, 4), seed=7), np.float64)
    u = np.random.default_rng(1).random((S, M, B, D))
    a = O.composer_generate(x, p, S, u)
    b = O.composer_generate(x, p, S, u)
    assert a.shape == (B, S, D, M) and np.array_equal(a, b) and set(np.unique(a)) <= {0., 1.}


def test_variable_lengths_drop_padded_rows():
    """utils/sequences.py:6-37: rows t >= lengths[b] are removed in b-major order; the values of the padded frames
    after a sequence's end cannot change any kept row (the recurrence is causal); full lengths = plain reshape."""
    B, T, D, M, H = 3, 5, 5, 2, 4
    lengths = np.array([5, 2, 3])
    x = O.synthetic_pianoroll(B, T, D, M, density=0.3).astype(np.float64)
    p = O.cast_params(O.init_composer_params(D, M, H, (6, 4), seed=7), np.float64)
    rows = O.flatten_valid_rows(lengths, T)
    np.testing.assert_array_equal(rows, [0, 1, 2, 3, 4, 5, 6, 10, 11, 12])
    full = O.composer_forward(x, p)
    out = O.composer_forward(x, p, lengths=lengths)
    np.testing.assert_array_equal(out['nll'], full['nll'][rows])
    np.testing.assert_allclose(out['loss'], full['nll'][rows].mean(), rtol=1e-13)
    x2 = x.copy()
    x2[1, 2:] = 1 - x2[1, 2:]                    # garbage after the end of sequence 1
    x2[2, 3:] = 1 - x2[2, 3:]
    out2 = O.composer_forward(x2, p, lengths=lengths)
    np.testing.assert_array_equal(out2['nll'], out['nll'])
    same = O.composer_forward(x, p, lengths=np.full(B, T))
    np.testing.assert_array_equal(same['nll'], full['nll'])
    tp = R.to_torch(p, torch.float64)
    loss, nll = R.composer_loss(torch.tensor(x), tp, lengths=lengths)
    np.testing.assert_allclose(float(loss), out['loss'], rtol=1e-12)
    np.testing.assert_allclose(nll.numpy(), out['nll'], rtol=1e-12)


# ----------------------------------------------------------------------------- pins against INDEPENDENT implementations
class _FixedUniforms:
    """Stands in for the numpy RandomState scikit-learn's RBM draws from: hands out the supplied uniform tensors."""

    def __init__(self, tensors):
        self.tensors = list(tensors)

    def uniform(self, size=None):
        u = self.tensors.pop(0)
        assert u.shape == tuple(size)
        return u


def test_rbm_oracle_equals_scikit_learn_bernoulli_rbm():
    """The oracle's RBM conditionals, strict-`<` Bernoulli sampling, k-step Gibbs chain and free energy (common/rbm.py:148-263,
    337-387) against scikit-learn's BernoulliRBM, an implementation that shares no code with the oracle."""
    from sklearn.neural_network import BernoulliRBM
    rng = np.random.default_rng(42)
    N, D, H, k = 13, 20, 9, 4
    W, bh, bv = rng.standard_normal((D, H)) * 0.7, rng.standard_normal((1, H)) * 0.4, rng.standard_normal((1, D)) * 0.4
    v0 = (rng.random((N, D)) < 0.3).astype(np.float64)
    uh, uv = rng.random((k, N, H)), rng.random((k, N, D))
    sk = BernoulliRBM(n_components=H)
    sk.components_, sk.intercept_hidden_, sk.intercept_visible_ = W.T.copy(), bh[0].copy(), bv[0].copy()
    np.testing.assert_allclose(O.rbm_cond_prob_h(v0, W, bh), sk._mean_hiddens(v0), rtol=1e-13)
    np.testing.assert_allclose(O.rbm_free_energy(v0, W, bh, bv), sk._free_energy(v0), rtol=1e-13)
    feed = _FixedUniforms([t for s in range(k) for t in (uh[s], uv[s])])
    v = v0
    for _ in range(k):
        h = sk._sample_hiddens(v, feed)
        v = sk._sample_visibles(h.astype(np.float64), feed).astype(np.float64)
    p_v, v_k = O.rbm_gibbs(v0, W, bh, bv, k, uh, uv)
    np.testing.assert_array_equal(v_k, v)
    assert not feed.tensors


def test_tf_adam_and_clip_equal_torch_optim_after_reparametrisation():
    """TF's Adam (epsilon added to the UNCORRECTED sqrt(v), train.py:64) is torch.optim.Adam with eps_t = eps / sqrt(1 -
    beta2^t); tf.clip_by_global_norm(5.) is torch's clip_grad_norm_ up to its 1e-6 guard. Ten steps on a quadratic."""
    import torch
    rng = np.random.default_rng(7)
    p0 = rng.standard_normal(50)
    A = rng.standard_normal((50, 50)) * 3
    grad = lambda p: A.T @ (A @ p)                                   # large gradients: the clip is active
    p, m, v = p0.copy(), np.zeros(50), np.zeros(50)
    tp = torch.tensor(p0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([tp], lr=0.01, betas=(0.9, 0.999), eps=1e-4)
    clipped = 0
    for t in range(1, 11):
        g = grad(p)
        (gc,), gn = O.clip_by_global_norm([g], 5.0)
        clipped += gn > 5.0
        p, m, v = O.tf_adam_step(p, gc, m, v, t)
        tp.grad = torch.tensor(grad(tp.detach().numpy()))
        torch.nn.utils.clip_grad_norm_([tp], 5.0)
        opt.param_groups[0]['eps'] = 1e-4 / np.sqrt(1 - 0.999 ** t)
        opt.step()
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=1e-6, atol=1e-9)
    assert clipped == 10


def test_cpu_philox_known_answers_and_counter_layouts():
    """oracle/philox.py against the Random123 known-answer vectors of Philox4x32-10, and the shapes / ranges / keying of
    the two counter layouts the kernels use."""
    from oracle.philox import gibbs_chain_uniforms, half_step_uniforms, philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32))
        assert tuple(int(x) for x in got) == want
    uh, uv = gibbs_chain_uniforms(11, 5, 6, 84, 256, 2)
    assert uh.shape == (2, 6, 256) and uv.shape == (2, 6, 84) and uh.dtype == np.float32
    assert 0.0 <= uh.min() and uh.max() < 1.0 and abs(float(uh.mean()) - 0.5) < 0.03
    # keyed by the global row: rows 2.. of a chain at offset 5 are rows 0.. of a chain at offset 7
    uh2, uv2 = gibbs_chain_uniforms(11, 7, 4, 84, 256, 2)
    assert np.array_equal(uh[:, 2:], uh2) and np.array_equal(uv[:, 2:], uv2)
    assert not np.array_equal(uh[0], uh[1]) and not np.array_equal(gibbs_chain_uniforms(12, 5, 6, 84, 256, 2)[0], uh)
    a = half_step_uniforms(3, 100, 4, 8)
    assert np.array_equal(a.reshape(-1)[8:], half_step_uniforms(3, 108, 3, 8).reshape(-1))     # offset = element index


def test_cpu_philox_nade_sampler_layout():
    from oracle.philox import nade_sample_uniforms, philox4x32_10
    u = nade_sample_uniforms(77, 3, 5, 4, 84)
    assert u.shape == (5, 4, 84) and u.dtype == np.float32 and 0 <= u.min() and u.max() < 1
    idx = (1 * 5 + 2) * 84 + 7                                             # row 1, track 2, dim 7 (row-major over tracks)
    w = philox4x32_10(np.array([idx, 0, 3, 0], np.uint32), np.array([77, 0], np.uint32))
    assert u[2, 1, 7] == np.float32(int(w[0]) >> 8) * np.float32(2.0 ** -24)
    assert not np.array_equal(u, nade_sample_uniforms(77, 4, 5, 4, 84))   # a new step draws new noise
    # a data-parallel shard (rows 2..3 of the global batch of 4) draws the global rows' noise
    assert np.array_equal(nade_sample_uniforms(77, 3, 5, 2, 84, row_map=(2, 4, 2)), u[:, 2:])


def test_cpu_philox_dropout_layout_is_shard_and_chunk_invariant():
    from oracle.philox import dropout_uniforms, philox4x32_10
    u = dropout_uniforms(9, 3, 6, 16)
    w = philox4x32_10(np.array([4 * 4 + 2, 0, 1, 0], np.uint32), np.array([9, 0], np.uint32))   # row 4, units 8..11, t=1
    assert u[1, 4, 9] == np.float32(int(w[1]) >> 8) * np.float32(2.0 ** -24)
    assert np.array_equal(dropout_uniforms(9, 3, 3, 16, row_map=(3, 6, 3)), u[:, 3:])
    assert np.array_equal(dropout_uniforms(9, 2, 6, 16, t_base=1), u[1:])


def test_rnn_rbm_generate_is_the_composition_of_its_steps():
    """oracle rnn_rbm_generate (rnn_estimator.py:271-323 for an RNN-RBM): the first generated frame is a k-step chain from
    the LAST INTRO FRAME under the biases of the intro's last LSTM output; later frames chain from the previous sample."""
    rng = np.random.default_rng(31)
    B, Ti, E, H, k, S = 3, 4, 12, 7, 2, 3
    p = O.init_rnn_rbm_params(I=E, D=E, H=H, R=(6, 5), seed=3)
    p = O.cast_params(p, np.float64)
    codes = (rng.random((B, Ti, E)) < 0.4).astype(np.float64)
    us = [(rng.random((k, B, H)), rng.random((k, B, E))) for _ in range(S)]
    out = O.rnn_rbm_generate(codes, p, k, S, us)
    assert out.shape == (B, S, E) and set(np.unique(out)) <= {0.0, 1.0}
    W, bh, bv = p['rbm']
    outs, state = O.rnn_scan(codes, p['lstm'])
    _, first = O.rbm_gibbs(codes[:, -1], W, bh + outs[:, -1] @ p['Wuh'], bv + outs[:, -1] @ p['Wuv'], k, *us[0])
    np.testing.assert_array_equal(out[:, 0], first)
    o, state = O.multi_rnn_step(first, state, p['lstm'])
    _, second = O.rbm_gibbs(first, W, bh + o @ p['Wuh'], bv + o @ p['Wuv'], k, *us[1])
    np.testing.assert_array_equal(out[:, 1], second)
    np.testing.assert_array_equal(out, O.rnn_rbm_generate(codes, p, k, S, us))          # pure function of the uniforms
