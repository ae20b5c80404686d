"""Composer / Jamming / Joint over DBN encoders against fixtures produced by the REFERENCE'S OWN CLASSES
(tests/golden/ref_dbn_modes.npz, written by tools/make_golden_ref.py::dbn_mode_cases from models/multinn/*.py,
encoders/dbn_encoder.py, common/{dbn,rbm,rnn,nade}.py, generators/rnn_{estimator,nade,multinade}.py -- unmodified -- on the
NumPy TF stand-in): the zero-padded inputs are encoded per track (Joint: stacked) to SAMPLED codes, RNN-NADE generators
model the codes, generated codes are decoded through the DBNs. All Bernoulli draws were logged in call order; the tests
replay them, so encode / sample / decode results are compared exactly.

CPU: the oracle composed the way the package's modes compose it reproduces the reference's losses and generated music.
GPU: the CUDA path through the public MultINN interface does. Nothing here reads /root/reference at run time."""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import torch_ref as R

Z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_dbn_modes.npz'))
TRACKS = ['Drums', 'Piano', 'Guitar', 'Bass', 'Strings']
MODES = ['composer', 'jamming', 'joint', 'feedback', 'feedback_rnn']      # feedback_rnn + DBN + NADE = BASELINE config C4
FEEDBACK = (40, 24)
X = Z['x'].astype(np.float64)
B, T, D, M = X.shape
E, H, RNN, ENC = 84, 128, (48, 32), (96, 84)
S = Z['composer/generate/music'].shape[1]
f64 = np.float64


def variables(mode):
    p = f'{mode}/var/'
    return {k[len(p):]: Z[k].astype(f64) for k in Z.files if k.startswith(p)}


def encoder_names(mode):
    return ['all'] if mode == 'joint' else TRACKS


def encoder_rbms(mode):
    v = variables(mode)
    return [[(v[f'multinn/dbn-encoder/{t}/dbn/rbm/{i}/W'], v[f'multinn/dbn-encoder/{t}/dbn/rbm/{i}/bh'],
              v[f'multinn/dbn-encoder/{t}/dbn/rbm/{i}/bv']) for i in range(2)] for t in encoder_names(mode)]


def lstm_layers(v, scope):
    return [(v[f'{scope}/multi_rnn_cell/cell_{i}/cudnn_compatible_lstm_cell/kernel'],
             v[f'{scope}/multi_rnn_cell/cell_{i}/cudnn_compatible_lstm_cell/bias']) for i in range(2)]


def generator_params(mode):
    v = variables(mode)
    if mode == 'composer':
        g = 'multinn/rnn-multinade'
        return dict(lstm=lstm_layers(v, g), dense=(v[f'{g}/dense/kernel'], v[f'{g}/dense/bias']),
                    nade=[(v[f'{g}/all/{t}/nade/w_enc'][:, 0, :], v[f'{g}/all/{t}/nade/w_dec'][:, :, 0]) for t in TRACKS])
    out = []
    for t in encoder_names(mode):
        g = f'multinn/rnn-nade/{t}'
        out.append(dict(lstm=lstm_layers(v, g), dense=(v[f'{g}/dense/kernel'], v[f'{g}/dense/bias']),
                        nade=(v[f'{g}/nade/w_enc'][:, 0, :], v[f'{g}/nade/w_dec'][:, :, 0])))
    return out


def feedback_params(mode):
    v = variables(mode)
    if mode == 'feedback':
        return [(v[f'multinn/feedback/dense{s}/kernel'], v[f'multinn/feedback/dense{s}/bias']) for s in ('', '_1')]
    return lstm_layers(v, 'multinn/feedback')


def _tt(a):
    if isinstance(a, (list, tuple)):
        return type(a)(_tt(b) for b in a)
    if isinstance(a, dict):
        return {k: _tt(b) for k, b in a.items()}
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64)


def encoder_inputs(mode):
    """Zero-padded encoder inputs [B,T+1,Dv]: per track (core/multi_encoder_nn.py:66-76) or stacked with feature
    d*M + m (multinn_joint.py:76-89)."""
    if mode == 'joint':
        return [np.concatenate([np.zeros((B, 1, D * M)), X.reshape(B, T, D * M)], axis=1)]
    return [np.concatenate([np.zeros((B, 1, D)), X[..., m]], axis=1) for m in range(M)]


def encode_draws(mode):
    """build() draws 4 arrays per encoder [encode 0, encode 1, reconstruct 1, reconstruct 0]; the codes the generators
    see come from the first two (rows b*(T+1) + t)."""
    n = len(encoder_names(mode))
    return [[Z[f'{mode}/eval/draw{4 * e}'].astype(f64), Z[f'{mode}/eval/draw{4 * e + 1}'].astype(f64)] for e in range(n)]


def oracle_codes(mode):
    codes = []
    for rows, rbms, us in zip(encoder_inputs(mode), encoder_rbms(mode), encode_draws(mode)):
        _, h = O.dbn_forward(rows.reshape(B * (T + 1), -1), rbms, us)
        codes.append(h.reshape(B, T + 1, -1))
    return codes


def oracle_loss(mode):
    codes = oracle_codes(mode)
    p = generator_params(mode)
    if mode == 'composer':
        stack = np.stack(codes, axis=3).reshape(B, T + 1, -1)                     # feature e*M + m
        outs, _ = O.rnn_scan(stack[:, :-1], p['lstm'])
        fc = O.dense(outs.reshape(B * T, -1), *p['dense'])
        be, bd = O.split_biases_multi(fc, M, H, E)
        tgt = stack[:, 1:].reshape(B * T, E, M)
        nll = np.stack([O.nade_log_prob(tgt[:, :, m], be[m], bd[m], *p['nade'][m])[0] for m in range(M)], 1)
        return np.array([nll.mean(0).mean()])
    if mode in ('feedback', 'feedback_rnn'):
        _, nll = R.feedback_loss(_tt(codes), _tt(p), _tt(feedback_params(mode)), 'dense' if mode == 'feedback' else 'rnn')
        return nll.mean(0).numpy()
    return np.array([O.rnn_nade_forward(c[:, :-1], c[:, 1:], q)['loss'] for c, q in zip(codes, p)])


def sampler_uniforms(mode):
    """The [n,B] sampler draws in call order -> u[S,G,B,E] with G generators' tracks: Composer draws step-major then
    track then dimension; Jamming runs each track's whole generation in turn; Joint has one track of E codes."""
    d = Z[f'{mode}/generate/sampler_draws'].astype(f64)
    if mode in ('composer', 'feedback', 'feedback_rnn'):
        return d.reshape(S, M, E, B).transpose(0, 1, 3, 2)
    if mode == 'jamming':
        return d.reshape(M, S, E, B).transpose(1, 0, 3, 2)
    return d.reshape(S, 1, E, B).transpose(0, 1, 3, 2)


def decode_draws(mode):
    """Per encoder [layer 1 -> 0, layer 0 -> visible] (common/dbn.py:158-180), rows b*S + s."""
    n = len(encoder_names(mode))
    return [[Z[f'{mode}/generate/decode_draw{2 * e}'].astype(f64), Z[f'{mode}/generate/decode_draw{2 * e + 1}'].astype(f64)]
            for e in range(n)]


def oracle_music(mode):
    codes, p, u = oracle_codes(mode), generator_params(mode), sampler_uniforms(mode)
    if mode == 'composer':
        h = O.multinade_generate(np.stack(codes, axis=3).reshape(B, T + 1, -1), p, S, u).reshape(B, S, E, M)
        h = [h[..., m] for m in range(M)]
    elif mode in ('feedback', 'feedback_rnn'):
        h = O.feedback_generate(codes, p, feedback_params(mode), 'dense' if mode == 'feedback' else 'rnn', S, u)
        h = [h[..., m] for m in range(M)]
    else:
        h = [O.multinade_generate(c, dict(lstm=q['lstm'], dense=q['dense'], nade=[q['nade']]), S, u[:, g:g + 1])
             for g, (c, q) in enumerate(zip(codes, p))]
    vis = [O.dbn_reconstruct(hm.reshape(B * S, E), rbms, us)[1] for hm, rbms, us in zip(h, encoder_rbms(mode), decode_draws(mode))]
    if mode == 'joint':
        return vis[0].reshape(B, S, D, M)
    return np.stack([v.reshape(B, S, D) for v in vis], axis=3)


# ----------------------------------------------------------------------------- CPU
@pytest.mark.parametrize('mode', MODES)
def test_oracle_loss_over_sampled_codes_matches_reference_code(mode):
    np.testing.assert_allclose(oracle_loss(mode), Z[f'{mode}/eval/loss'], rtol=1e-10)


@pytest.mark.parametrize('mode', MODES)
def test_oracle_generate_through_dbn_matches_reference_code(mode):
    np.testing.assert_array_equal(oracle_music(mode), Z[f'{mode}/generate/music'])


def test_draw_counts():
    """4 build draws per encoder + 2 per encoder for decoding the predictions; generation: S*tracks*E sampler draws."""
    for mode, n_enc in (('composer', 5), ('jamming', 5), ('joint', 1), ('feedback', 5), ('feedback_rnn', 5)):
        assert int(Z[f'{mode}/eval/n_draw']) == 6 * n_enc
        assert Z[f'{mode}/generate/sampler_draws'].shape == (S * E * (1 if mode == 'joint' else M), B)


# ----------------------------------------------------------------------------- GPU
def _model(mode, device='cuda'):
    from multinn_b200.multinn import MultINN, default_config, default_params
    from multinn_b200.utils.tf_import import load_tf_variables
    name = mode.replace('_', '-')
    kw = {'feedback': list(FEEDBACK)} if mode.startswith('feedback') else {}
    m = MultINN(default_config(), default_params(mode=name, encoder='DBN', encoder_hidden=list(ENC), generator='NADE',
                                                 num_hidden=H, num_hidden_rnn=RNN, keep_prob=1.0, **kw), name, device=device)
    v = variables(mode)
    load_tf_variables(m, {k: a for k, a in v.items() if 'dbn-encoder' not in k}, strict=True)
    load_tf_variables(m, {k: a for k, a in v.items() if 'dbn-encoder' in k}, which='encoders', strict=True)
    return m


def _time_major(u, steps):
    """Rows b*steps + t (the reference's flatten order) -> t*B + b (the device's staging order)."""
    return np.ascontiguousarray(u.reshape(B, steps, -1).transpose(1, 0, 2).reshape(steps * B, -1))


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


@pytest.mark.parametrize('mode', MODES)
def test_reference_named_dbn_variables_import_strictly(mode):
    m = _model(mode, device='cpu')
    esd = {k: v.numpy() for k, v in m.encoder_arena.state_dict().items()}
    rbms = encoder_rbms(mode)
    first = 'encoder/all' if mode == 'joint' else 'encoder/Drums'
    np.testing.assert_array_equal(esd[f'{first}/rbm_1/W'], rbms[0][1][0].astype(np.float32))
    np.testing.assert_array_equal(esd[f'{first}/rbm_0/bv'].reshape(-1), rbms[0][0][2].astype(np.float32).reshape(-1))


@pytest.mark.gpu
@pytest.mark.parametrize('mode', MODES)
def test_cuda_evaluate_over_dbn_codes_matches_reference_code(mode):
    m = _model(mode)
    u_enc = [[_cu(_time_major(a, T + 1)) for a in us] for us in encode_draws(mode)]
    out = m.evaluate(_cu(X), u_enc=u_enc if mode != 'joint' else u_enc[0])
    ref = Z[f'{mode}/eval/loss']
    np.testing.assert_allclose(float(out['batch/loss']), ref.mean(), rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('mode', MODES)
def test_cuda_generate_through_dbn_matches_reference_code(mode):
    m = _model(mode)
    u_enc = [[_cu(_time_major(a, T + 1)) for a in us] for us in encode_draws(mode)]
    u_dec = [[_cu(a) for a in us] for us in decode_draws(mode)]
    u = _cu(sampler_uniforms(mode))
    if mode == 'joint':
        u_enc, u_dec = u_enc[0], u_dec[0]
    music = m.generate(_cu(X), S, u=u, u_enc=u_enc, u_dec=u_dec)
    np.testing.assert_array_equal(music.cpu().numpy().astype(np.uint8), Z[f'{mode}/generate/music'])


# ----------------------------------------------------------------------------- Joint + RNN-RBM (BASELINE config C3)
# tests/golden/ref_joint_rbm.npz: the reference's own Joint mode with its RNN-RBM generator, run with the two outside
# stand-ins its HEAD needs to start at all (tools/make_golden_ref.py::joint_rbm_case).
J = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_joint_rbm.npz'))
JX = J['x'].astype(np.float64)
JB, JT = JX.shape[:2]
JK, JH, JS = int(J['k']), 64, J['generate/music'].shape[1]


def _jvars():
    return {k[4:]: J[k].astype(f64) for k in J.files if k.startswith('var/')}


def _joint_oracle_params():
    v = _jvars()
    enc = [(v[f'multinn/dbn-encoder/all/dbn/rbm/{i}/W'], v[f'multinn/dbn-encoder/all/dbn/rbm/{i}/bh'],
            v[f'multinn/dbn-encoder/all/dbn/rbm/{i}/bv']) for i in range(2)]
    g = 'multinn/rnn-rbm'
    lstm = [(v[f'{g}/{g}/all/multi_rnn_cell/cell_{i}/cudnn_compatible_lstm_cell/kernel'],
             v[f'{g}/{g}/all/multi_rnn_cell/cell_{i}/cudnn_compatible_lstm_cell/bias']) for i in range(2)]
    gen = dict(lstm=lstm, rbm=(v[f'{g}/all/rbm/W'], v[f'{g}/all/rbm/bh'], v[f'{g}/all/rbm/bv']),
               Wuh=v[f'{g}/all/Wuh'], Wuv=v[f'{g}/all/Wuv'])
    return enc, gen


def _jdraw(section, i):
    return J[f'{section}/draw{i}'].astype(f64)


def _joint_codes(enc):
    """Sampled codes [B,T+1,E] of the stacked zero-padded inputs from the build()'s first two draws (rows b*(T+1) + t)."""
    pad = np.concatenate([np.zeros((JB, 1, D * M)), JX.reshape(JB, JT, D * M)], axis=1)
    _, h = O.dbn_forward(pad.reshape(JB * (JT + 1), -1), enc, [_jdraw('eval', 0), _jdraw('eval', 1)])
    return h.reshape(JB, JT + 1, -1)


def test_oracle_joint_rnn_rbm_matches_reference_code():
    """Free-energy cost, chain samples and the monitored log-likelihood of the teacher-forced pass (rnn_rbm.py:94-119),
    then generation (multinn_joint.py:188-215), against the reference's own classes."""
    enc, gen = _joint_oracle_params()
    codes = _joint_codes(enc)
    uh = np.stack([_jdraw('eval', 4 + 2 * i) for i in range(JK)])            # [k, N, H], rows n = b*T + t
    uv = np.stack([_jdraw('eval', 5 + 2 * i) for i in range(JK)])
    r = O.rnn_rbm_forward(codes[:, :-1], codes[:, 1:], gen, JK, uh, uv)
    np.testing.assert_allclose(r['loss'], J['eval/loss'], rtol=1e-10)
    np.testing.assert_array_equal(r['sample'], J['eval/sample'])
    tgt = codes[:, 1:].reshape(JB * JT, -1)
    ll = -(tgt * np.log(r['cond_p'] + 1e-7) + (1 - tgt) * np.log(1 - r['cond_p'] + 1e-7)).sum(1).mean()
    np.testing.assert_allclose(ll, J['eval/log_likelihood'], rtol=1e-10)
    # generation: per step k x [h, v] draws [B, .], then the decode (rows b*S + s); the intro codes are build()'s
    us = [(np.stack([_jdraw('generate', s * 2 * JK + 2 * i) for i in range(JK)]),
           np.stack([_jdraw('generate', s * 2 * JK + 2 * i + 1) for i in range(JK)])) for s in range(JS)]
    u_dec = [_jdraw('generate', JS * 2 * JK), _jdraw('generate', JS * 2 * JK + 1)]
    h = O.rnn_rbm_generate(codes, gen, JK, JS, us)
    _, vis = O.dbn_reconstruct(h.reshape(JB * JS, -1), enc, u_dec)
    np.testing.assert_array_equal(vis.reshape(JB, JS, D, M), J['generate/music'])


def _joint_model(device='cuda'):
    from multinn_b200.multinn import MultINN, default_config, default_params
    from multinn_b200.utils.tf_import import load_tf_variables
    m = MultINN(default_config(), default_params(mode='joint', encoder='DBN', encoder_hidden=list(ENC), generator='RBM',
                                                 num_hidden=JH, num_hidden_rnn=RNN, keep_prob=1.0), 'joint', device=device)
    core = m._model
    core._generator._k = core._generator.rbm._k = JK
    v = _jvars()
    load_tf_variables(m, {k: a for k, a in v.items() if 'dbn-encoder' not in k}, strict=True)
    load_tf_variables(m, {k: a for k, a in v.items() if 'dbn-encoder' in k}, which='encoders', strict=True)
    return m


def test_reference_named_rnn_rbm_variables_import_strictly():
    m = _joint_model(device='cpu')
    sd = {k: v.numpy() for k, v in m.arena.state_dict().items()}
    _, gen = _joint_oracle_params()
    np.testing.assert_array_equal(sd['generator/Wuv'], gen['Wuv'].astype(np.float32))
    np.testing.assert_array_equal(sd['generator/rbm/bh'].reshape(-1), gen['rbm'][1].astype(np.float32).reshape(-1))
    np.testing.assert_array_equal(sd['generator/rnn/cell_1/kernel'], gen['lstm'][1][0].astype(np.float32))


@pytest.mark.gpu
def test_cuda_joint_rnn_rbm_matches_reference_code():
    """C3 through the public interface: evaluate() (free-energy cost, chain samples) and generate() with the reference's
    logged draws -- cost within 1e-4, samples and generated music bit for bit."""
    m = _joint_model()
    tm = lambda a, steps: _cu(_time_major(a, steps))              # reference rows b*steps + t -> device rows t*B + b
    u_enc = [tm(_jdraw('eval', 0), JT + 1), tm(_jdraw('eval', 1), JT + 1)]
    uh = torch.stack([tm(_jdraw('eval', 4 + 2 * i), JT) for i in range(JK)])
    uv = torch.stack([tm(_jdraw('eval', 5 + 2 * i), JT) for i in range(JK)])
    out = m.evaluate(_cu(JX), u_enc=u_enc, u_gibbs=(uh, uv))
    np.testing.assert_allclose(float(out['batch/loss']), float(J['eval/loss']), rtol=1e-4, atol=1e-4)
    got = out['sample'].cpu().numpy().reshape(JT, JB, -1).transpose(1, 0, 2).reshape(JB * JT, -1)
    np.testing.assert_array_equal(got.astype(np.uint8), J['eval/sample'])
    np.testing.assert_allclose(float(out['log_likelihood']), float(J['eval/log_likelihood']), rtol=2e-4)
    us = [(torch.stack([_cu(_jdraw('generate', s * 2 * JK + 2 * i)) for i in range(JK)]),
           torch.stack([_cu(_jdraw('generate', s * 2 * JK + 2 * i + 1)) for i in range(JK)])) for s in range(JS)]
    u_dec = [_cu(_jdraw('generate', JS * 2 * JK)), _cu(_jdraw('generate', JS * 2 * JK + 1))]
    music = m.generate(_cu(JX), JS, u=us, u_enc=u_enc, u_dec=u_dec)
    np.testing.assert_array_equal(music.cpu().numpy().astype(np.uint8), J['generate/music'])
