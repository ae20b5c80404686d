"""GPU parity: RBM / DBN primitives, the Joint (DBN + LSTM-RBM) mode and the Feedback / Feedback-RNN modes against the
CPU oracle with the same weights and the same uniform-noise tensors."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import torch_ref as R

pytestmark = pytest.mark.gpu
f64 = np.float64


def make(mode, **kw):
    from multinn_b200.multinn import MultINN, default_config, default_params
    kw.setdefault('keep_prob', 1.0)
    return MultINN(default_config(), default_params(mode=mode, **kw), mode)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def sd_np(arena):
    return {k: v.numpy() for k, v in arena.state_dict().items()}


def rnn_nade_params(sd, prefix, L):
    return dict(lstm=[(sd[f'{prefix}/rnn/cell_{l}/kernel'], sd[f'{prefix}/rnn/cell_{l}/bias']) for l in range(L)],
                dense=(sd[f'{prefix}/dense/kernel'], sd[f'{prefix}/dense/bias']),
                nade=(sd[f'{prefix}/nade/w_enc'][0], sd[f'{prefix}/nade/w_dec'][0]))


def dbn_params(sd, prefix, L):
    return [(sd[f'{prefix}/rbm_{i}/W'], sd[f'{prefix}/rbm_{i}/bh'], sd[f'{prefix}/rbm_{i}/bv']) for i in range(L)]


# ----------------------------------------------------------------------------- RBM / DBN primitives
def test_rbm_gibbs_free_energy_and_cd():
    from multinn_b200.common.rbm import RBM
    from multinn_b200.params import ParamArena
    rng = np.random.default_rng(0)
    N, D, H, k = 70, 84, 256, 3
    arena = ParamArena()
    rbm = RBM(D, H, k=k, arena=arena, name='rbm')
    arena.finalize('cuda', seed=1)
    arena.load('rbm/bh', rng.standard_normal((1, H)) * 0.2)
    arena.load('rbm/bv', rng.standard_normal((1, D)) * 0.2)
    W, bh, bv = (sd_np(arena)[n].astype(f64) for n in ('rbm/W', 'rbm/bh', 'rbm/bv'))
    v = (rng.random((N, D)) < 0.3).astype(np.float32)
    bh_t = (rng.standard_normal((N, H)) * 0.3).astype(np.float32)
    bv_t = (rng.standard_normal((N, D)) * 0.3).astype(np.float32)
    uh, uv = rng.random((k, N, H), dtype=np.float32), rng.random((k, N, D), dtype=np.float32)
    p_v, vk = rbm.sample(cu(v), cu(bh_t), cu(bv_t), u=(cu(uh), cu(uv)))
    rp, rv = O.rbm_gibbs(v.astype(f64), W, bh_t.astype(f64), bv_t.astype(f64), k, uh.astype(f64), uv.astype(f64))
    np.testing.assert_array_equal(vk.cpu().numpy(), rv)
    np.testing.assert_allclose(p_v.cpu().numpy(), rp, rtol=2e-5)
    p0, v0 = rbm.sample(cu(v), k=0)
    assert torch.equal(v0, cu(v))
    cost, fe = rbm.free_energy_cost(cu(v), vk)
    ref = O.rbm_free_energy_cost_mean(v.astype(f64), rv, W, bh, bv)
    assert abs(float(cost) - ref) < 1e-4 * max(1.0, abs(ref))
    assert abs(float(fe) - O.rbm_free_energy(v.astype(f64), W, bh, bv).mean()) < 1e-3
    # gradient of the cost (internal biases) vs autograd
    t = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    Wt, bht, bvt = t(W), t(bh), t(bv)
    R.rbm_free_energy_cost_mean(torch.tensor(v, dtype=torch.float64), torch.tensor(rv), Wt, bht, bvt).backward()
    arena.grad.zero_()
    rbm.free_energy_cost_backward(cu(v), vk)
    for p, g in ((rbm.W, Wt.grad), (rbm.bh, bht.grad), (rbm.bv, bvt.grad)):
        assert float((p.grad.cpu().double() - g).norm() / g.norm()) < 1e-4
    # CD-k update
    u = dict(uh=uh, uv=uv, uh0=rng.random((N, H), dtype=np.float32), uhk=rng.random((N, H), dtype=np.float32))
    dW, dbv, dbh = O.rbm_cd_update(v.astype(f64), W, bh, bv, k, 0.05, uh.astype(f64), uv.astype(f64),
                                   u['uh0'].astype(f64), u['uhk'].astype(f64))
    rbm.train(cu(v), 0.05, u={k_: cu(a) for k_, a in u.items()})
    new = sd_np(arena)
    np.testing.assert_allclose(new['rbm/W'], W + dW, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(new['rbm/bv'], bv + dbv, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(new['rbm/bh'], bh + dbh, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('N,D,H,k,rowbias', [(70, 84, 256, 3, True), (33, 84, 168, 2, False), (64, 168, 84, 2, True),
                                             (5, 20, 64, 10, True), (1, 84, 256, 1, False), (258, 84, 128, 4, True)])
def test_fused_gibbs_chain_equals_oracle_and_half_step_path(N, D, H, k, rowbias):
    """mnn_rbm_gibbs (one launch per chain, W + W^T in shared memory) against the fp64 oracle chain (rbm.py:192-231) and
    against the GEMM + half-step path, same uniforms: samples bit-exact, last-step probabilities to 2e-5."""
    from multinn_b200 import _lib, ops
    from multinn_b200.common.rbm import RBM
    from multinn_b200.params import ParamArena
    rng = np.random.default_rng(N * 1000 + D + H)
    arena = ParamArena()
    rbm = RBM(D, H, k=k, arena=arena, name='rbm')
    arena.finalize('cuda', seed=7)
    arena.load('rbm/bh', rng.standard_normal((1, H)) * 0.2)
    arena.load('rbm/bv', rng.standard_normal((1, D)) * 0.2)
    W, bh, bv = (sd_np(arena)[n].astype(f64) for n in ('rbm/W', 'rbm/bh', 'rbm/bv'))
    v = (rng.random((N, D)) < 0.1).astype(np.float32)
    v[N // 2] = 0.0                                                    # an all-zero row (every input dim skipped)
    bh_t = (rng.standard_normal((N, H)) * 0.3).astype(np.float32) if rowbias else None
    bv_t = (rng.standard_normal((N, D)) * 0.3).astype(np.float32) if rowbias else None
    bh_o = bh if bh_t is None else bh_t.astype(f64)
    bv_o = bv if bv_t is None else bv_t.astype(f64)
    for _ in range(200):     # uniforms whose every comparison along the fp64 chain has a margin fp32 rounding cannot cross
        uh, uv = rng.random((k, N, H), dtype=np.float32), rng.random((k, N, D), dtype=np.float32)
        vv, margin = v.astype(f64), 1.0
        for s_ in range(k):
            ph = O.rbm_cond_prob_h(vv, W, bh_o)
            hh = (uh[s_] < ph).astype(f64)
            pv = O.rbm_cond_prob_v(hh, W, bv_o)
            vv = (uv[s_] < pv).astype(f64)
            margin = min(margin, float(np.abs(uh[s_] - ph).min()), float(np.abs(uv[s_] - pv).min()))
        if margin > 2e-6:
            break
    assert margin > 2e-6
    assert ops.rbm_gibbs_supported(cu(v), rbm.W.data, None, None, (cu(uh), cu(uv)))
    args = (cu(v), None if bh_t is None else cu(bh_t), None if bv_t is None else cu(bv_t))
    default = ops.GIBBS_MODE
    assert default == 'auto' and ops.gibbs_use_fused(N) and not ops.gibbs_use_fused(ops.GIBBS_FUSED_MAX_ROWS + 1)
    launches = _lib.lib.mnn_launch_count()
    p_f, v_f = rbm.sample(*args, u=(cu(uh), cu(uv)))
    assert _lib.lib.mnn_launch_count() - launches == 1               # the whole chain was one kernel
    ops.GIBBS_MODE = 'gemm'
    try:
        p_g, v_g = rbm.sample(*args, u=(cu(uh), cu(uv)))
        assert _lib.lib.mnn_launch_count() - launches >= 1 + 4 * k   # 2k GEMMs + 2k half-steps
    finally:
        ops.GIBBS_MODE = default
    rp, rv = O.rbm_gibbs(v.astype(f64), W, bh_o, bv_o, k, uh.astype(f64), uv.astype(f64))
    np.testing.assert_array_equal(v_f.cpu().numpy(), rv)
    np.testing.assert_array_equal(v_g.cpu().numpy(), rv)
    np.testing.assert_allclose(p_f.cpu().numpy(), rp, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(p_f.cpu().numpy(), p_g.cpu().numpy(), rtol=2e-5, atol=1e-7)
    # direct call: h_k output, a strided v0 view and the k-prefix of longer uniform tensors
    wide = torch.zeros(N, D + 12, device='cuda')
    wide[:, 4:4 + D] = cu(v)
    h_k = torch.empty(N, H, device='cuda')
    v_k = torch.empty(N, D, device='cuda')
    ops.rbm_gibbs(wide[:, 4:4 + D], rbm.W.data, args[1] if rowbias else rbm.bh.data, args[2] if rowbias else rbm.bv.data,
                  k, v_k=v_k, h_k=h_k, u=(cu(uh), cu(uv)))
    assert torch.equal(v_k, v_f)
    vv = v.astype(f64)
    for s in range(k):                                                 # oracle half-steps up to h_k
        _, hh = O.rbm_forward(vv, W, bh_o, uh[s].astype(f64))
        _, vv = O.rbm_reconstruct(hh, W, bv_o, uv[s].astype(f64))
    np.testing.assert_array_equal(h_k.cpu().numpy(), hh)


def test_fused_gibbs_philox_statistics_and_determinism():
    from multinn_b200 import ops
    from multinn_b200.common.rbm import RBM
    from multinn_b200.params import ParamArena
    arena = ParamArena()
    rbm = RBM(84, 256, k=2, arena=arena, name='rbm')
    arena.finalize('cuda', seed=2)
    rbm.W.data.zero_()                              # zero weights and biases: every conditional is exactly 0.5
    v = torch.zeros(4098, 84, device='cuda')
    p1, v1 = rbm.sample(v)
    assert torch.all(p1 == 0.5) and set(np.unique(v1.cpu().numpy())) <= {0.0, 1.0}
    assert abs(float(v1.mean()) - 0.5) < 0.005
    cols = v1.mean(0)
    assert float((cols - 0.5).abs().max()) < 0.05   # no stuck column group
    _, v2 = rbm.sample(v)
    assert not torch.equal(v1, v2)                  # fresh noise on every call (quirk Q12)
    outs = []
    for _ in range(2):                              # same (seed, offset): the same chain
        vk, hk = torch.empty(4098, 84, device='cuda'), torch.empty(4098, 256, device='cuda')
        ops.rbm_gibbs(v, rbm.W.data, rbm.bh.data, rbm.bv.data, 2, v_k=vk, h_k=hk, seed=11, offset=5)
        outs.append((vk, hk))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert abs(float(outs[0][1].mean()) - 0.5) < 0.005
    # keyed by the global row: rows [100:200) of a chain at offset 5 == rows [0:100) at offset 105
    vk = torch.empty(100, 84, device='cuda')
    ops.rbm_gibbs(v[:100], rbm.W.data, rbm.bh.data, rbm.bv.data, 2, v_k=vk, seed=11, offset=105)
    assert torch.equal(vk, outs[0][0][100:200])


def test_fused_gibbs_philox_stream_equals_cpu_philox():
    """The in-kernel generator of mnn_rbm_gibbs, not only its statistics: the CPU Philox4x32-10 with the kernel's counter
    layout (oracle/philox.py) reproduces the uniforms, so the Philox-mode chain must equal the oracle chain bit for bit."""
    from multinn_b200 import ops
    from oracle.philox import gibbs_chain_uniforms
    rng = np.random.default_rng(17)
    N, D, H, k, offset = 37, 84, 256, 3, 1000
    W = (rng.standard_normal((D, H)) * 0.1).astype(np.float32)
    bh = (rng.standard_normal((N, H)) * 0.3).astype(np.float32)
    bv = (rng.standard_normal((N, D)) * 0.3).astype(np.float32)
    v0 = (rng.random((N, D)) < 0.2).astype(np.float32)
    for seed in range(100, 400):                      # a seed whose every comparison has a margin fp32 cannot cross
        uh, uv = gibbs_chain_uniforms(seed, offset, N, D, H, k)
        vv, margin = v0.astype(f64), 1.0
        for s_ in range(k):
            ph = O.rbm_cond_prob_h(vv, W.astype(f64), bh.astype(f64))
            hh = (uh[s_] < ph).astype(f64)
            pv = O.rbm_cond_prob_v(hh, W.astype(f64), bv.astype(f64))
            vv = (uv[s_] < pv).astype(f64)
            margin = min(margin, float(np.abs(uh[s_] - ph).min()), float(np.abs(uv[s_] - pv).min()))
        if margin > 2e-6:
            break
    assert margin > 2e-6
    v_k, h_k = torch.empty(N, D, device='cuda'), torch.empty(N, H, device='cuda')
    ops.rbm_gibbs(cu(v0), cu(W), cu(bh), cu(bv), k, v_k=v_k, h_k=h_k, seed=seed, offset=offset)
    np.testing.assert_array_equal(v_k.cpu().numpy(), vv)
    np.testing.assert_array_equal(h_k.cpu().numpy(), hh)


def test_half_step_philox_stream_equals_cpu_philox():
    from multinn_b200 import ops
    from oracle.philox import half_step_uniforms
    rng = np.random.default_rng(18)
    N, C, offset = 29, 168, 12345
    pre = (rng.standard_normal((N, C))).astype(np.float32)
    p_ref = O.sigmoid(pre.astype(f64))
    for seed in range(7, 200):
        u = half_step_uniforms(seed, offset, N, C)
        if float(np.abs(u - p_ref).min()) > 2e-6:
            break
    s = torch.empty(N, C, device='cuda')
    ops.bias_sigmoid_sample(cu(pre), s=s, use_philox=True, seed=seed, offset=offset)
    np.testing.assert_array_equal(s.cpu().numpy(), (u < p_ref).astype(np.float32))


def test_rbm_philox_half_step_statistics():
    from multinn_b200.common.rbm import RBM
    from multinn_b200.params import ParamArena
    arena = ParamArena()
    rbm = RBM(84, 256, arena=arena, name='rbm')
    arena.finalize('cuda', seed=2)
    v = torch.zeros(4096, 84, device='cuda')
    p, h = rbm.forward(v)                      # zero input, zero bias -> p = 0.5
    assert set(np.unique(h.cpu().numpy())) <= {0.0, 1.0}
    assert abs(float(h.mean()) - 0.5) < 0.005
    _, h2 = rbm.forward(v)
    assert not torch.equal(h, h2)              # fresh noise on every call (quirk Q12)


def test_dbn_encoder_encode_decode_and_layerwise_cd():
    from multinn_b200.encoders.dbn_encoder import DBNEncoder
    from multinn_b200.params import ParamArena
    rng = np.random.default_rng(3)
    arena = ParamArena()
    enc = DBNEncoder(84, [168, 84], arena=arena, name='enc')
    arena.finalize('cuda', seed=4)
    rbms = [tuple(a.astype(f64) for a in r) for r in dbn_params(sd_np(arena), 'enc', 2)]
    N = 50
    x = (rng.random((N, 84)) < 0.2).astype(np.float32)
    us = [rng.random((N, 168), dtype=np.float32), rng.random((N, 84), dtype=np.float32)]
    p_h, h = enc.encode(cu(x), u=[cu(a) for a in us])
    rp, rh = O.dbn_forward(x.astype(f64), rbms, [a.astype(f64) for a in us])
    np.testing.assert_array_equal(h.cpu().numpy(), rh)
    np.testing.assert_allclose(p_h.cpu().numpy(), rp, rtol=2e-5)
    ud = [rng.random((N, 168), dtype=np.float32), rng.random((N, 84), dtype=np.float32)]
    p_v, v = enc.decode(h, u=[cu(a) for a in ud])
    rpv, rv = O.dbn_reconstruct(rh, rbms, [a.astype(f64) for a in ud])
    np.testing.assert_array_equal(v.cpu().numpy(), rv)
    # layer-1 CD-1 on the sampled layer-0 codes
    k = 1
    u = dict(lower=[cu(us[0])], cd={n: cu(a) for n, a in dict(
        uh=rng.random((k, N, 84), dtype=np.float32), uv=rng.random((k, N, 168), dtype=np.float32),
        uh0=rng.random((N, 84), dtype=np.float32), uhk=rng.random((N, 84), dtype=np.float32)).items()})
    _, h0 = O.rbm_forward(x.astype(f64), rbms[0][0], rbms[0][1], us[0].astype(f64))
    W1, bh1, bv1 = rbms[1]
    c = {n: a.cpu().numpy().astype(f64) for n, a in u['cd'].items()}
    dW, dbv, dbh = O.rbm_cd_update(h0, W1, bh1, bv1, k, 0.1, c['uh'], c['uv'], c['uh0'], c['uhk'])
    enc.train(cu(x), 0.1, layer=1, u=u)
    np.testing.assert_allclose(sd_np(arena)['enc/rbm_1/W'], W1 + dW, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd_np(arena)['enc/rbm_0/W'], rbms[0][0], rtol=0, atol=0)   # lower layer frozen


# ----------------------------------------------------------------------------- Joint: DBN encoder + LSTM-RBM (config C3)
def _joint_uniforms(rng, B, T, k, H=64):
    N1, N = (T + 1) * B, T * B
    return dict(u_enc=[rng.random((N1, 168), dtype=np.float32), rng.random((N1, 84), dtype=np.float32)],
                u_gibbs=(rng.random((k, N, H), dtype=np.float32), rng.random((k, N, 84), dtype=np.float32)))


def test_joint_dbn_rnn_rbm_parity_and_gradients():
    B, T, H, Rn, k = 6, 7, 64, (48, 32), 10
    model = make('joint', encoder='DBN', encoder_hidden=[168, 84], generator='RBM', num_hidden=H, num_hidden_rnn=Rn)
    rng = np.random.default_rng(5)
    core = model._model
    core.arena.load('generator/rbm/bh', rng.standard_normal((1, H)) * 0.1)
    core.arena.load('generator/rbm/bv', rng.standard_normal((1, 84)) * 0.1)
    sd, esd = sd_np(core.arena), sd_np(core.encoder_arena)
    x = O.synthetic_pianoroll(B, T, seed=8, density=0.1)
    u = _joint_uniforms(rng, B, T, k, H)
    out = model.evaluate(cu(x), u_enc=[cu(a) for a in u['u_enc']], u_gibbs=tuple(cu(a) for a in u['u_gibbs']))
    # oracle (time-major row order n' = t*B + b to share the uniforms)
    rbms = [tuple(a.astype(f64) for a in r) for r in dbn_params(esd, 'encoder/all', 2)]
    inp, _ = O.composer_inputs_targets(x)
    pad = np.concatenate([inp, x.reshape(B, T, -1)[:, -1:]], axis=1)             # [B,T+1,420]
    flat_tm = pad.transpose(1, 0, 2).reshape((T + 1) * B, -1).astype(f64)
    _, codes = O.dbn_forward(flat_tm, rbms, [a.astype(f64) for a in u['u_enc']])
    np.testing.assert_array_equal(out['codes'].cpu().numpy().reshape((T + 1) * B, -1), codes)
    codes_bm = codes.reshape(T + 1, B, -1).transpose(1, 0, 2)                    # [B,T+1,84]
    p = dict(lstm=[(sd[f'generator/rnn/cell_{l}/kernel'].astype(f64), sd[f'generator/rnn/cell_{l}/bias'].astype(f64))
                   for l in range(2)],
             rbm=tuple(sd[f'generator/rbm/{n}'].astype(f64) for n in ('W', 'bh', 'bv')),
             Wuh=sd['generator/Wuh'].astype(f64), Wuv=sd['generator/Wuv'].astype(f64))
    to_bm = lambda a, C: a.reshape(a.shape[0], T, B, C).transpose(0, 2, 1, 3).reshape(a.shape[0], B * T, C)
    ref = O.rnn_rbm_forward(codes_bm[:, :-1], codes_bm[:, 1:], p, k, to_bm(u['u_gibbs'][0].astype(f64), H),
                            to_bm(u['u_gibbs'][1].astype(f64), 84))
    got_sample = out['sample'].cpu().numpy().reshape(T, B, -1).transpose(1, 0, 2).reshape(B * T, -1)
    np.testing.assert_array_equal(got_sample, ref['sample'])
    assert abs(float(out['batch/loss']) - ref['loss']) < 1e-4 * max(1.0, abs(ref['loss']))       # generator's own loss: no /M
    # the monitored log-likelihood (common/rbm.py:121-129): per-row sum of tf.losses.log_loss(targets, cond_probs), eps 1e-7
    tgt_bm, p_bm = codes_bm[:, 1:].reshape(B * T, -1), ref['cond_p']
    nll_ref = -(tgt_bm * np.log(p_bm + 1e-7) + (1 - tgt_bm) * np.log(1 - p_bm + 1e-7)).sum(1)
    np.testing.assert_allclose(out['nll'].cpu().numpy()[:, 0], nll_ref, rtol=2e-4, atol=1e-4)
    # training step: gradient reaches W, bh, bv only (quirk Q3); LSTM / Wuh / Wuv stay untouched
    before = core.arena.state_dict()
    step = model.train_generators('sgd', 0.1)
    step(cu(x), u_enc=[cu(a) for a in u['u_enc']], u_gibbs=tuple(cu(a) for a in u['u_gibbs']))
    after = core.arena.state_dict()
    t = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    Wt, bht, bvt = (t(a) for a in p['rbm'])
    tgt = torch.tensor(codes_bm[:, 1:].reshape(B * T, -1))
    R.rbm_free_energy_cost_mean(tgt, torch.tensor(ref['sample']), Wt, bht, bvt).backward()
    gn = float(torch.sqrt(Wt.grad.pow(2).sum() + bht.grad.pow(2).sum() + bvt.grad.pow(2).sum()))
    scale = 5.0 / max(gn, 5.0)
    np.testing.assert_allclose(after['generator/rbm/W'].numpy(), p['rbm'][0] - 0.1 * scale * Wt.grad.numpy(),
                               rtol=1e-4, atol=1e-6)
    for n in ('generator/Wuh', 'generator/Wuv', 'generator/rnn/cell_0/kernel'):
        assert torch.equal(before[n], after[n])


def test_pretrain_generators_updates_only_the_rbm_module():
    """multinn_joint.py / rnn_rbm.py:299-322: pre-training = CD-k on the generator's RBM over the flattened input codes;
    RNN-NADE generators have nothing to pre-train (rnn_nade.py:320-326)."""
    model = make('joint', encoder='DBN', encoder_hidden=[168, 84], generator='RBM', num_hidden=64, num_hidden_rnn=(32,))
    x = cu(O.synthetic_pianoroll(3, 5, seed=2, density=0.1))
    before = model._model.arena.state_dict()
    step = model.pretrain_generators(None, 0.05)
    assert step(x) == 1
    after = model._model.arena.state_dict()
    assert not torch.equal(before['generator/rbm/W'], after['generator/rbm/W'])
    assert torch.isfinite(after['generator/rbm/W']).all()
    for n in before:
        if not n.startswith('generator/rbm/'):
            assert torch.equal(before[n], after[n]), n
    comp = make('composer', num_hidden=128, num_hidden_rnn=(32,))
    b0 = comp._model.arena.flat.clone()
    assert comp.pretrain_generators(None, 0.05)(x) == 0
    assert torch.equal(b0, comp._model.arena.flat)


def test_joint_with_nade_generator_nll_parity():
    """Joint mode with an RNN-NADE generator over the DBN codes (one "track" of 84 code bits): per-row NLL against the
    oracle on the same sampled codes; `batch/loss` is the generator's own mean NLL (generator.py:196-201; the /M of
    multinn_joint.py:182-184 applies to the encoder-level `global` metrics only)."""
    B, T = 5, 6
    model = make('joint', encoder='DBN', encoder_hidden=[168, 84], generator='NADE', num_hidden=128, num_hidden_rnn=(48, 32))
    core = model._model
    rng = np.random.default_rng(33)
    sd, esd = sd_np(core.arena), sd_np(core.encoder_arena)
    x = O.synthetic_pianoroll(B, T, seed=14, density=0.1)
    N1 = (T + 1) * B
    u_enc = [rng.random((N1, 168), dtype=np.float32), rng.random((N1, 84), dtype=np.float32)]
    out = model.evaluate(cu(x), u_enc=[cu(a) for a in u_enc])
    rbms = [tuple(a.astype(f64) for a in r) for r in dbn_params(esd, 'encoder/all', 2)]
    inp, _ = O.composer_inputs_targets(x)
    pad = np.concatenate([inp, x.reshape(B, T, -1)[:, -1:]], axis=1)
    _, codes = O.dbn_forward(pad.transpose(1, 0, 2).reshape(N1, -1).astype(f64), rbms, [a.astype(f64) for a in u_enc])
    codes = codes.reshape(T + 1, B, -1).transpose(1, 0, 2)                                   # [B,T+1,84]
    ref = O.rnn_nade_forward(codes[:, :-1], codes[:, 1:], O.cast_params(rnn_nade_params(sd, 'generator', 2), f64))
    np.testing.assert_allclose(out['nll'].cpu().numpy()[:, 0], ref['nll'], rtol=1e-4)
    assert abs(float(out['batch/loss']) - ref['nll'].mean()) < 1e-4 * ref['nll'].mean()


def test_joint_generate_bit_exact_against_oracle():
    """multinn_joint.py:188-215 end to end with supplied uniforms: DBN-encode the intro, RNN-RBM generation (k-step chain
    from the previous frame, LSTM step, new biases), DBN-decode: the generated piano-rolls equal the oracle's bit for bit."""
    B, Ti, S, H, Rn, k = 4, 5, 3, 64, (48, 32), 4
    model = make('joint', encoder='DBN', encoder_hidden=[168, 84], generator='RBM', num_hidden=H, num_hidden_rnn=Rn)
    core = model._model
    core._generator._k = core._generator.rbm._k = k
    rng = np.random.default_rng(21)
    core.arena.load('generator/rbm/bh', rng.standard_normal((1, H)) * 0.1)
    core.arena.load('generator/rbm/bv', rng.standard_normal((1, 84)) * 0.1)
    sd, esd = sd_np(core.arena), sd_np(core.encoder_arena)
    x = O.synthetic_pianoroll(B, Ti, seed=4, density=0.1)
    N1 = (Ti + 1) * B
    u_enc = [rng.random((N1, 168), dtype=np.float32), rng.random((N1, 84), dtype=np.float32)]
    us = [(rng.random((k, B, H), dtype=np.float32), rng.random((k, B, 84), dtype=np.float32)) for _ in range(S)]
    u_dec = [rng.random((B * S, 168), dtype=np.float32), rng.random((B * S, 420), dtype=np.float32)]
    got = model.generate(cu(x), S, u=[(cu(a), cu(b)) for a, b in us], u_enc=[cu(a) for a in u_enc],
                         u_dec=[cu(a) for a in u_dec]).cpu().numpy()
    rbms = [tuple(a.astype(f64) for a in r) for r in dbn_params(esd, 'encoder/all', 2)]
    p = dict(lstm=[(sd[f'generator/rnn/cell_{l}/kernel'].astype(f64), sd[f'generator/rnn/cell_{l}/bias'].astype(f64))
                   for l in range(2)],
             rbm=tuple(sd[f'generator/rbm/{n}'].astype(f64) for n in ('W', 'bh', 'bv')),
             Wuh=sd['generator/Wuh'].astype(f64), Wuv=sd['generator/Wuv'].astype(f64))
    ref = O.joint_generate(x.astype(f64), rbms, p, k, S, [a.astype(f64) for a in u_enc],
                           [(a.astype(f64), b.astype(f64)) for a, b in us], [a.astype(f64) for a in u_dec])
    assert got.shape == ref.shape == (B, S, 84, 5)
    np.testing.assert_array_equal(got, ref)


def test_joint_generate_runs_and_is_binary():
    model = make('joint', encoder='DBN', encoder_hidden=[168, 84], generator='RBM', num_hidden=64, num_hidden_rnn=(32,))
    x = cu(O.synthetic_pianoroll(3, 5, seed=2))
    s = model.generate(x, 4)
    assert s.shape == (3, 4, 84, 5) and set(np.unique(s.cpu().numpy())) <= {0.0, 1.0}


def test_joint_rbm_fit_loop_monitors_generator_log_likelihood(tmp_path):
    """train.py:153-282 for the reference's Joint config (DBN + RNN-RBM): the epoch loop needs the generator's
    `log_likelihood` (common/rbm.py:121-129) from evaluate(); it must run, stay finite and checkpoint."""
    from multinn_b200.utils import training as U
    rng = np.random.default_rng(5)
    X = (rng.random((6, 8, 84, 5)) < 0.08).astype(np.uint8)
    lengths = np.full(6, 8)
    model = make('joint', encoder='DBN', encoder_hidden=[168, 84], generator='RBM', num_hidden=64, num_hidden_rnn=(32,))
    cfg = {'batch_size': 3, 'piece_size': 8, 'learning_rate': 0.01, 'epochs': 2, 'early_stopping': 5}
    stats, hist = U.fit(model, (X, lengths), (X[:4], lengths[:4]), cfg, checkpoint_path=str(tmp_path / 'best.pt'))
    assert stats.epoch == 2 and len(hist) == 2
    assert all(np.isfinite(h['valid_log_likelihood']) and h['valid_log_likelihood'] > 0 for h in hist)


# ----------------------------------------------------------------------------- Composer / Jamming over DBN encodings
def _track_codes(x, esd, tracks, u_enc, B, T):
    """Oracle: per-track DBN codes of the zero-padded inputs (core/multi_encoder_nn.py:66-115), [M][B,T+1,E]."""
    codes = []
    for m, t in enumerate(tracks):
        rbms = [tuple(a.astype(f64) for a in r) for r in dbn_params(esd, f'encoder/{t}', 2)]
        pad = np.concatenate([np.zeros((B, 1, 84)), x[..., m]], axis=1)                      # [B,T+1,D]
        rows = pad.transpose(1, 0, 2).reshape((T + 1) * B, 84).astype(f64)                   # time-major like the device
        _, h = O.dbn_forward(rows, rbms, [a.astype(f64) for a in u_enc[m]])
        codes.append(h.reshape(T + 1, B, -1).transpose(1, 0, 2))
    return codes


@pytest.mark.parametrize('mode', ['composer', 'jamming'])
def test_composer_and_jamming_over_dbn_encodings(mode):
    """core/multi_encoder_nn.py:98-115: the generators of Composer / Jamming run over the per-track encoders' SAMPLED codes
    (inputs codes[:, :-1], targets codes[:, 1:]); with the same uniforms the per-row NLL equals the oracle's on the same
    codes, one training step runs, and generation decodes back to [B,S,84,5] binary piano-rolls."""
    B, T = 4, 5
    model = make(mode, encoder='DBN', encoder_hidden=[168, 84], generator='NADE', num_hidden=128, num_hidden_rnn=(48, 32))
    core = model._model
    rng = np.random.default_rng(41)
    sd, esd = sd_np(core.arena), sd_np(core.encoder_arena)
    x = O.synthetic_pianoroll(B, T, seed=15, density=0.1)
    N1 = (T + 1) * B
    u_enc = [[rng.random((N1, 168), dtype=np.float32), rng.random((N1, 84), dtype=np.float32)] for _ in range(5)]
    out = model.evaluate(cu(x), u_enc=[[cu(a) for a in um] for um in u_enc])
    codes = _track_codes(x, esd, core.tracks, u_enc, B, T)
    if mode == 'composer':
        stack = np.stack(codes, axis=3).reshape(B, T + 1, -1)                                # feature e*M + m
        p = O.cast_params(dict(lstm=[(sd[f'generator/rnn/cell_{l}/kernel'], sd[f'generator/rnn/cell_{l}/bias'])
                                     for l in range(2)],
                               dense=(sd['generator/dense/kernel'], sd['generator/dense/bias']),
                               nade=[(sd['generator/nade/w_enc'][m], sd['generator/nade/w_dec'][m]) for m in range(5)]), f64)
        outs, _ = O.rnn_scan(stack[:, :-1], p['lstm'])
        fc = O.dense(outs.reshape(B * T, -1), *p['dense'])
        be, bd = O.split_biases_multi(fc, 5, 128, 84)
        tgt = stack[:, 1:].reshape(B * T, 84, 5)
        ref = np.stack([O.nade_log_prob(tgt[:, :, m], be[m], bd[m], *p['nade'][m])[0] for m in range(5)], 1)
    else:
        ref = np.stack([O.rnn_nade_forward(codes[m][:, :-1], codes[m][:, 1:],
                                           O.cast_params(rnn_nade_params(sd, f'generator/{t}', 2), f64))['nll']
                        for m, t in enumerate(core.tracks)], 1)
    np.testing.assert_allclose(out['nll'].cpu().numpy(), ref, rtol=1e-4)
    step = model.train_generators('adam', 0.01)
    l0 = float(step(cu(x), u_enc=[[cu(a) for a in um] for um in u_enc]))
    assert abs(l0 - ref.mean(0).mean()) < 1e-4 * ref.mean()
    s = model.generate(cu(x), 3)
    assert s.shape == (B, 3, 84, 5) and set(np.unique(s.cpu().numpy())) <= {0.0, 1.0}


# ----------------------------------------------------------------------------- Feedback / Feedback-RNN (config C4)
def _fb_case(mode, encoder):
    kw = dict(num_hidden=128, num_hidden_rnn=(48, 32), feedback=[40, 24])
    if encoder == 'DBN':
        kw.update(encoder='DBN', encoder_hidden=[168, 84])
    return make(mode, **kw)


@pytest.mark.parametrize("mode,encoder", [('feedback', 'Pass'), ('feedback-rnn', 'Pass'), ('feedback-rnn', 'DBN')])
def test_feedback_modes_nll_gradients_and_generation(mode, encoder):
    B, T, S = 4, 6, 5
    model = _fb_case(mode, encoder)
    core = model._model
    rng = np.random.default_rng(9)
    sd, esd = sd_np(core.arena), sd_np(core.encoder_arena)
    x = O.synthetic_pianoroll(B, T, seed=12, density=0.1)
    kind = 'dense' if mode == 'feedback' else 'rnn'
    M = 5
    # per-track encodings of the zero-padded inputs (time-major rows for the DBN uniforms)
    pad = np.concatenate([np.zeros((B, 1, 84, M), np.float32), x], axis=1)        # [B,T+1,84,M]
    u_enc = None
    if encoder == 'DBN':
        u_enc = [[rng.random(((T + 1) * B, 168), dtype=np.float32), rng.random(((T + 1) * B, 84), dtype=np.float32)]
                 for _ in range(M)]
        xe = []
        for m, t in enumerate(model.tracks):
            rbms = [tuple(a.astype(f64) for a in r) for r in dbn_params(esd, f'encoder/{t}', 2)]
            flat = pad[..., m].transpose(1, 0, 2).reshape((T + 1) * B, 84).astype(f64)
            _, h = O.dbn_forward(flat, rbms, [a.astype(f64) for a in u_enc[m]])
            xe.append(h.reshape(T + 1, B, 84).transpose(1, 0, 2))
    else:
        xe = [pad[..., m].astype(f64) for m in range(M)]
    gp = [O.cast_params(rnn_nade_params(sd, f'generator/{t}', 2), f64) for t in model.tracks]
    if kind == 'dense':
        fbp = [(sd[f'feedback/dense_{l}/kernel'].astype(f64), sd[f'feedback/dense_{l}/bias'].astype(f64)) for l in range(2)]
    else:
        fbp = [(sd[f'feedback/rnn/cell_{l}/kernel'].astype(f64), sd[f'feedback/rnn/cell_{l}/bias'].astype(f64))
               for l in range(2)]
    tt = lambda tree: R.to_torch(tree, torch.float64, requires_grad=True)
    gpt, fbt = [tt(p) for p in gp], tt(fbp)
    loss, nll = R.feedback_loss([torch.tensor(a) for a in xe], gpt, fbt, kind)
    cu_enc = None if u_enc is None else [[cu(a) for a in uu] for uu in u_enc]
    out = model.evaluate(cu(x), u_enc=cu_enc)
    np.testing.assert_allclose(out['nll'].cpu().numpy(), nll.detach().numpy(), rtol=1e-4)
    # gradients of every generator + the feedback module
    leaves = R.flat_params(gpt) + R.flat_params(fbt)
    grads = torch.autograd.grad(loss, leaves)
    core.arena.grad.zero_()
    l = core._forward_backward(core._check_x(cu(x), None), keep=1.0, u_drop=None, seed=0, u_enc=cu_enc)
    assert abs(float(l) - float(loss)) / float(loss) < 1e-5
    named = core.arena.named()
    fb_names = [f'feedback/dense_{i}/{n}' for i in range(2) for n in ('kernel', 'bias')] if kind == 'dense' else \
        [f'feedback/rnn/cell_{i}/{n}' for i in range(2) for n in ('kernel', 'bias')]
    for name, g in zip(fb_names, grads[-4:]):
        got = named[name].grad.cpu().double()
        assert float((got - g).norm() / g.norm()) < 2e-4, name
    g0 = grads[0]                                           # first generator's layer-0 kernel
    got = named[f'generator/{model.tracks[0]}/rnn/cell_0/kernel'].grad.cpu().double()
    assert float((got - g0).norm() / g0.norm()) < 2e-4
    # generation: sampled encodings bit-exact (Pass encoders decode = identity)
    if encoder == 'Pass':
        u = rng.random((S, M, B, 84), dtype=np.float32)
        got = model.generate(cu(x), S, u=cu(u)).cpu().numpy()
        ref = O.feedback_generate(xe, gp, fbp, kind, S, u.astype(f64))
        np.testing.assert_array_equal(got, ref)


def test_feedback_training_reduces_loss():
    model = _fb_case('feedback-rnn', 'Pass')
    step = model.train_generators('adam', 0.01)
    x = cu(O.synthetic_pianoroll(4, 6, seed=1))
    l0 = float(step(x, keep=1.0))
    for _ in range(6):
        l1 = float(step(x, keep=1.0))
    assert l1 < l0


def test_train_encoders_step_matches_oracle_cd_and_driver_runs(tmp_path):
    """Encoder pre-training (train_encoders.py, multi_encoder_nn.py:155-195): the rows each encoder trains on are the
    zero-padded inputs cut by the unpadded lengths; one `step` = the oracle's CD-k update per track; the init_ops set
    the visible bias from the data mean; the layer-wise driver lowers the reconstruction log-loss."""
    from multinn_b200.utils import training as U
    rng = np.random.default_rng(11)
    B, T, M, k = 3, 5, 5, 1
    model = make('jamming', encoder='DBN', encoder_hidden=[24, 12], num_hidden=128, num_hidden_rnn=(16,))
    core = model._model
    x = O.synthetic_pianoroll(B, T, seed=3, density=0.2)
    lengths = np.array([5, 3, 4])
    rows = core._encoder_rows(cu(x), torch.from_numpy(lengths))
    pad = np.concatenate([np.zeros((B, 1, 84, M), np.float32), x], 1)              # [B,T+1,D,M]
    for m in range(M):
        ref_rows = np.concatenate([pad[b, :lengths[b], :, m] for b in range(B)])     # b-major; order is irrelevant
        got = rows[m].cpu().numpy()
        assert got.shape == ref_rows.shape
        assert np.array_equal(np.sort(got.sum(1)), np.sort(ref_rows.sum(1))) and got.sum() == ref_rows.sum()
    full = core._encoder_rows(cu(x), None)
    assert full[0].shape == (T * B, 84) and float(full[0][:B].abs().sum()) == 0.0      # zero frame first
    # one CD-1 step on layer 0, supplied uniforms, against the oracle update on the same rows
    N = rows[0].shape[0]
    sd0 = sd_np(core.encoder_arena)
    u = []
    for m in range(M):
        cd = dict(uh=rng.random((k, N, 24), dtype=np.float32), uv=rng.random((k, N, 84), dtype=np.float32),
                  uh0=rng.random((N, 24), dtype=np.float32), uhk=rng.random((N, 24), dtype=np.float32))
        met = dict(lower=[], up=rng.random((N, 24), dtype=np.float32), down=rng.random((N, 84), dtype=np.float32))
        u.append(dict(train=dict(lower=[], cd={n: cu(a) for n, a in cd.items()}),
                      metrics=dict(lower=[], up=cu(met['up']), down=cu(met['down'])), _cd=cd))
    init, step = model.train_encoders(None, 0.05, layer=0)
    out = step(cu(x), lengths=torch.from_numpy(lengths), u=u)
    assert np.isfinite(float(out['batch/loss'])) and float(out['log_likelihood']) > 0
    sd1 = sd_np(core.encoder_arena)
    for m, t in enumerate(model.tracks):
        W, bh, bv = (sd0[f'encoder/{t}/rbm_0/{n}'].astype(f64) for n in ('W', 'bh', 'bv'))
        c = {n: a.astype(f64) for n, a in u[m]['_cd'].items()}
        dW, dbv, dbh = O.rbm_cd_update(rows[m].cpu().numpy().astype(f64), W, bh, bv, k, 0.05, c['uh'], c['uv'], c['uh0'], c['uhk'])
        np.testing.assert_allclose(sd1[f'encoder/{t}/rbm_0/W'], W + dW, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(sd1[f'encoder/{t}/rbm_0/bv'], bv + dbv.reshape(bv.shape), rtol=1e-4, atol=1e-6)
        np.testing.assert_array_equal(sd1[f'encoder/{t}/rbm_1/W'], sd0[f'encoder/{t}/rbm_1/W'])
    init(cu(x), lengths=torch.from_numpy(lengths))
    sd2 = sd_np(core.encoder_arena)
    np.testing.assert_allclose(sd2['encoder/Drums/rbm_0/bv'].reshape(-1),
                               O.rbm_visible_bias_init(rows[0].cpu().numpy().astype(f64)).reshape(-1), rtol=1e-4, atol=1e-5)
    # the driver: two layers, a few epochs each
    X = (rng.random((8, 6, 84, 5)) < 0.1).astype(np.uint8)
    L = np.array([6, 6, 4, 6, 2, 6, 5, 6])
    cfg = {'batch_size': 4, 'piece_size': 6, 'learning_rate': 0.1, 'epochs': 8, 'early_stopping': 20}
    for layer in (0, 1):
        stats, hist = U.fit_encoders(model, (X, L), (X[:4], L[:4]), cfg, layer=layer,
                                     checkpoint_path=str(tmp_path / f'enc{layer}.pt'))
        assert stats.epoch == 8 and stats.steps == 16 and len(hist) == 8
        assert all(np.isfinite(h['valid_log_likelihood']) and np.isfinite(h['loss']) for h in hist)
        if layer == 0:          # layer 1 trains on freshly sampled binary codes: too noisy at this size to assert on
            assert min(h['valid_log_likelihood'] for h in hist[1:]) < hist[0]['valid_log_likelihood']


@pytest.mark.parametrize("mode", ['jamming', 'feedback', 'feedback-rnn'])
def test_chunk_pipeline_in_multi_generator_modes(mode):
    """T = 64 switches the per-generator time-chunk pipeline on (five generators one after another, `need_dx` through
    the aux-stream dx GEMMs in the feedback modes): loss and every gradient equal the unpipelined path."""
    from multinn_b200.common.rnn import RNN
    B, T = 8, 64
    x = cu(O.synthetic_pianoroll(B, T, seed=21, density=0.07))
    results = []
    saved = RNN.PIPE_MAX_BATCH
    try:
        for pipe in (saved if saved > 0 else 512, 0):
            RNN.PIPE_MAX_BATCH = pipe
            model = _fb_case(mode, 'Pass') if mode != 'jamming' else make('jamming', num_hidden=128, num_hidden_rnn=(32, 16))
            core = model._model
            core.arena.grad.zero_()
            loss = core._forward_backward(core._check_x(x, None), keep=1.0, u_drop=None, seed=0)
            torch.cuda.synchronize()
            results.append((float(loss), core.arena.grad.clone()))
    finally:
        RNN.PIPE_MAX_BATCH = saved
    (l1, g1), (l0, g0) = results
    assert abs(l1 - l0) / l0 < 1e-6
    assert float((g1 - g0).norm() / g0.norm()) < 1e-5


@pytest.mark.parametrize("mode", ['feedback', 'feedback-rnn'])
def test_feedback_modes_variable_lengths(mode):
    """Variable `lengths` reach the generators only (multinn_feedback.py:93-94): kept-row NLL, loss and the gradients of
    the feedback module and of a generator against the oracle."""
    B, T, M = 4, 7, 5
    lengths = np.array([7, 3, 5, 1])
    model = _fb_case(mode, 'Pass')
    core = model._model
    sd = sd_np(core.arena)
    x = O.synthetic_pianoroll(B, T, seed=14, density=0.1)
    kind = 'dense' if mode == 'feedback' else 'rnn'
    pad = np.concatenate([np.zeros((B, 1, 84, M), np.float32), x], axis=1)
    xe = [pad[..., m].astype(f64) for m in range(M)]
    gp = [O.cast_params(rnn_nade_params(sd, f'generator/{t}', 2), f64) for t in model.tracks]
    names = [f'feedback/dense_{i}/{n}' for i in range(2) for n in ('kernel', 'bias')] if kind == 'dense' else \
        [f'feedback/rnn/cell_{i}/{n}' for i in range(2) for n in ('kernel', 'bias')]
    fbp = [(sd[names[2 * l]].astype(f64), sd[names[2 * l + 1]].astype(f64)) for l in range(2)]
    tt = lambda tree: R.to_torch(tree, torch.float64, requires_grad=True)
    gpt, fbt = [tt(p) for p in gp], tt(fbp)
    loss, nll = R.feedback_loss([torch.tensor(a) for a in xe], gpt, fbt, kind, lengths=lengths)
    out = model.evaluate(cu(x), lengths=torch.from_numpy(lengths))
    assert out['nll'].shape == (int(lengths.sum()), M)
    np.testing.assert_allclose(out['nll'].cpu().numpy(), nll.detach().numpy(), rtol=1e-4)
    leaves = R.flat_params(gpt) + R.flat_params(fbt)
    grads = torch.autograd.grad(loss, leaves)
    core.arena.grad.zero_()
    l = core._forward_backward(core._check_x(cu(x), lengths), keep=1.0, u_drop=None, seed=0, lengths=lengths)
    assert abs(float(l) - float(loss)) / float(loss) < 1e-5
    named = core.arena.named()
    for name, g in zip(names, grads[-4:]):
        got = named[name].grad.cpu().double()
        assert float((got - g).norm() / g.norm()) < 2e-4, name
    got = named[f'generator/{model.tracks[0]}/rnn/cell_0/kernel'].grad.cpu().double()
    assert float((got - grads[0]).norm() / grads[0].norm()) < 2e-4
