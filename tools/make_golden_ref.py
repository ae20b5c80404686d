"""Generates tests/golden/ref_primitives.npz by EXECUTING THE REFERENCE'S OWN MODULES
(/root/reference/multinn/models/common/{nade,rbm,dbn}.py, utils/sequences.py, models/generators/rnn_multinade.py,
metrics/statistical.py -- imported unmodified) on the NumPy-backed `tensorflow` / `tensorflow_probability` stand-ins in
tests/tf_stub/. Loop order, transposes, reshapes, eps placement, Gibbs-chain structure, CD-k formula, bias split and
flatten order are therefore the reference's code; only the per-op semantics are NumPy's (float64).

Runs in the build container only (it reads /root/reference); the .npz travels. Weights are drawn here from a seeded
Generator and ASSIGNED into the reference objects' variables, and every Bernoulli draw consumes an injected uniform
array, so the oracle and the CUDA path can be fed exactly the same numbers.

  python tools/make_golden_ref.py        # rewrites tests/golden/ref_primitives.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/multinn'
sys.path.insert(0, os.path.join(ROOT, 'tests', 'tf_stub'))
sys.path.insert(0, REF)

import tensorflow as tf  # noqa: E402  (the stub)
import tensorflow_probability as tfp  # noqa: E402  (the stub)

assert tf.__version__.endswith('numpy-stub'), 'the real TensorFlow must not shadow tests/tf_stub'

from models.common.dbn import DBN  # noqa: E402
from models.common.nade import NADE  # noqa: E402
from models.common.rbm import RBM  # noqa: E402
from utils.sequences import flatten_maybe_padded_sequences  # noqa: E402

A = np.asarray
out = {}


def r32(a):
    """Round to float32-representable values (kept in float64): the CUDA path is fed the same numbers exactly."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


class _R32:
    """Generator whose float draws are float32-representable."""

    def __init__(self, seed):
        self._g = np.random.default_rng(seed)

    def random(self, shape):
        return r32(self._g.random(shape, dtype=np.float32))

    def standard_normal(self, shape):
        return r32(self._g.standard_normal(shape))


def put(prefix, **kw):
    for k, v in kw.items():
        out[f'{prefix}/{k}'] = A(v)


def nade_cases():
    """NADE.log_prob (nade.py:155-229) and NADE.sample (:231-308): external biases, the shapes the generators use."""
    for name, (N, D, H, density, seed) in {'d05': (24, 84, 128, 0.05, 1), 'd50': (16, 84, 256, 0.5, 2),
                                           'd100': (5, 84, 128, 1.0, 3), 'd0': (4, 84, 128, 0.0, 4),
                                           'small': (9, 20, 12, 0.3, 5)}.items():
        rng = _R32(seed)
        nade = NADE(D, H, internal_bias=False, name=f'nade_{name}')
        w_enc = r32(rng.standard_normal((D, 1, H)) / np.sqrt(D))
        w_dec = r32(rng.standard_normal((D, H, 1)) / np.sqrt(D))
        nade.w_enc.assign(w_enc)
        nade.w_dec.assign(w_dec)
        x = (rng.random((N, D)) < density).astype(np.float64)
        b_enc = rng.standard_normal((N, H))
        b_dec = r32(rng.standard_normal((N, D)) - 1.5)
        nll, cond_p = nade.log_prob(tf.constant(x), tf.constant(b_enc), tf.constant(b_dec))
        u = rng.random((N, D))
        tfp.push_uniforms([u[:, i:i + 1] for i in range(D)])        # one [N,1] draw per dimension, in loop order
        sample, sample_nll = nade.sample(tf.constant(b_enc), tf.constant(b_dec), temperature=1.)
        assert tfp.pending() == 0
        thr, thr_nll = nade.sample(tf.constant(b_enc), tf.constant(b_dec), temperature=None)
        put(f'nade/{name}', x=x, b_enc=b_enc, b_dec=b_dec, w_enc=w_enc[:, 0, :], w_dec=w_dec[:, :, 0], nll=nll,
            cond_p=cond_p, u=u, sample=sample, sample_nll=sample_nll, sample_threshold=thr, threshold_nll=thr_nll)
    # a [1,H] / [1,D] bias row is tiled over the batch (nade.py:184-187)
    rng = _R32(6)
    nade = NADE(20, 12, name='nade_tile')
    nade.w_enc.assign(r32(A(nade.w_enc)))
    nade.w_dec.assign(r32(A(nade.w_dec)))
    x = (rng.random((7, 20)) < 0.4).astype(np.float64)
    b_enc, b_dec = rng.standard_normal((1, 12)), rng.standard_normal((1, 20))
    nll, cond_p = nade.log_prob(tf.constant(x), tf.constant(b_enc), tf.constant(b_dec))
    put('nade/tile', x=x, b_enc=b_enc, b_dec=b_dec, w_enc=A(nade.w_enc)[:, 0, :], w_dec=A(nade.w_dec)[:, :, 0], nll=nll,
        cond_p=cond_p)


def rbm_cases():
    """RBM.forward / reconstruct / sample / free_energy_cost / visible_bias_init_ops / _cd_update (rbm.py:148-373)."""
    for name, (N, D, H, k, per_row, seed) in {'gen': (12, 84, 64, 3, True, 11), 'enc': (10, 20, 8, 2, False, 12),
                                              'k10': (6, 84, 256, 10, True, 13)}.items():
        rng = _R32(seed)
        rbm = RBM(D, H, k=k, name=f'rbm_{name}')
        W = r32(A(rbm.W))
        rbm.W.assign(W)
        bh0 = r32(rng.standard_normal((1, H)) * 0.2)
        bv0 = r32(rng.standard_normal((1, D)) * 0.2)
        rbm.bh.assign(bh0)
        rbm.bv.assign(bv0)
        v = (rng.random((N, D)) < 0.15).astype(np.float64)
        bh = r32(rng.standard_normal((N, H)) * 0.5) if per_row else None
        bv = r32(rng.standard_normal((N, D)) * 0.5 - 1.0) if per_row else None
        tb = lambda a: None if a is None else tf.constant(a)
        uh, uv = rng.random((k, N, H)), rng.random((k, N, D))
        tfp.push_uniforms([uh[0]])
        p_h, h = rbm.forward(tf.constant(v), tb(bh))
        tfp.push_uniforms([uv[0]])
        p_v1, v1 = rbm.reconstruct(h, tb(bv))
        tfp.push_uniforms([a for s in range(k) for a in (uh[s], uv[s])])       # while_loop: forward then reconstruct
        p_vk, v_k = rbm.sample(tf.constant(v), tb(bh), tb(bv), k=k)
        assert tfp.pending() == 0
        tgt = (rng.random((N, D)) < 0.15).astype(np.float64)
        cost, free_energy = rbm.free_energy_cost(tf.constant(tgt), v_k)          # internal biases (quirk Q3), [N,N] (Q4)
        metrics, _, _ = rbm.build_metrics(tf.constant(tgt), v_k, cond_probs=p_vk)
        put(f'rbm/{name}', W=W, bh0=bh0, bv0=bv0, v=v, bh=np.zeros(0) if bh is None else bh,
            bv=np.zeros(0) if bv is None else bv, uh=uh, uv=uv, k=k, p_h=p_h, h=h, p_v1=p_v1, v1=v1, p_vk=p_vk, v_k=v_k,
            target=tgt, cost_shape=A(A(cost).shape), cost_mean=A(cost).mean(), free_energy=free_energy,
            batch_loss=metrics['batch/loss'], log_likelihood=metrics['log_likelihood'])
        # CD-k update with the internal biases (rbm.py:299-335); _cd_update directly: train() would also run the
        # visible-bias init op, which in graph mode is a separate fetch (train_encoders.py:106-109)
        uh0, uhk = rng.random((N, H)), rng.random((N, H))
        tfp.push_uniforms([a for s in range(k) for a in (uh[s], uv[s])] + [uh0, uhk])
        lr = float(np.float32(0.05))
        rbm._cd_update(tf.constant(v), lr)
        assert tfp.pending() == 0
        put(f'rbm/{name}/cd', lr=lr, uh0=uh0, uhk=uhk, W=A(rbm.W), bh=A(rbm.bh), bv=A(rbm.bv))
    rng = _R32(14)
    rbm = RBM(20, 8, k=1, name='rbm_bvinit')
    v = (rng.random((50, 20)) < 0.3).astype(np.float64)
    v[:, 3] = 0.0                                                    # a never-on unit: log(1e-6 + 0)
    rbm.visible_bias_init_ops(tf.constant(v))
    put('rbm/bvinit', v=v, bv=A(rbm.bv))


def dbn_cases():
    """DBN.forward / reconstruct (dbn.py:136-180): chained sampled half-steps, reconstruct in reverse layer order."""
    rng = _R32(21)
    N, D, hidden = 9, 84, [40, 24]
    dbn = DBN(D, hidden, k=1, name='dbn')
    for i, rbm in enumerate(dbn.rbm_layers):
        rbm.W.assign(r32(A(rbm.W)))
        rbm.bh.assign(r32(rng.standard_normal(A(rbm.bh).shape) * 0.2))
        rbm.bv.assign(r32(rng.standard_normal(A(rbm.bv).shape) * 0.2))
    v = (rng.random((N, D)) < 0.1).astype(np.float64)
    u_fwd = [rng.random((N, h)) for h in hidden]
    tfp.push_uniforms(u_fwd)
    p_h, h = dbn.forward(tf.constant(v))
    u_rec = [rng.random((N, hidden[0])), rng.random((N, D))]            # last layer first
    tfp.push_uniforms(u_rec)
    p_v, v_rec = dbn.reconstruct(h)
    assert tfp.pending() == 0
    kw = {}
    for i, rbm in enumerate(dbn.rbm_layers):
        kw.update({f'W{i}': A(rbm.W), f'bh{i}': A(rbm.bh), f'bv{i}': A(rbm.bv), f'u_fwd{i}': u_fwd[i], f'u_rec{i}': u_rec[i]})
    put('dbn', v=v, p_h=p_h, h=h, p_v=p_v, v_rec=v_rec, **kw)


def sequence_cases():
    """flatten_maybe_padded_sequences (utils/sequences.py:6-37): b-major rows, padded steps removed."""
    rng = _R32(31)
    t = rng.random((4, 5, 3))
    put('seq/full', tensor=t, lengths=A([5, 5, 5, 5]),
        flat=flatten_maybe_padded_sequences(tf.constant(t), tf.constant(A([5, 5, 5, 5]))))
    put('seq/none', tensor=t, flat=flatten_maybe_padded_sequences(tf.constant(t), None))
    lengths = A([5, 2, 4, 1])
    put('seq/ragged', tensor=t, lengths=lengths, flat=flatten_maybe_padded_sequences(tf.constant(t), tf.constant(lengths)))


def multinade_cases():
    """RnnMultiNADE._build_biases (rnn_multinade.py:231-256: [M*H | M*D] split, M chunks each), the per-track log_prob
    loop (:281-288), build_metrics (:154-210: mean over tracks of the per-track batch means; metrics/statistical.py:34)
    and sample_single (:292-317: stack axis=2 -> feature d*M + m). The object is created without running the RNN /
    Dense construction (tf.contrib cells are not part of the stub); the methods under test use only the attributes set here."""
    from models.generators.rnn_multinade import RnnMultiNADE
    from models.generators.rnn_estimator import RnnEstimatorStateTuple
    rng = _R32(41)
    M, D, H, N = 5, 84, 128, 11
    tracks = ['Drums', 'Piano', 'Guitar', 'Bass', 'Strings']
    g = object.__new__(RnnMultiNADE)
    g._name, g._tracks, g._num_dims, g._num_hidden, g._internal_bias = 'rnn-multinade', tracks, D, [H], False
    g._num_output = M * D
    g._track_name = 'all'
    g._nades = []
    w = {}
    for m in range(M):
        nade = NADE(D, H, name=f'nade_{tracks[m]}')
        nade.w_enc.assign(r32(rng.standard_normal((D, 1, H)) / np.sqrt(D)))
        nade.w_dec.assign(r32(rng.standard_normal((D, H, 1)) / np.sqrt(D)))
        g._nades.append(nade)
        w[f'w_enc{m}'], w[f'w_dec{m}'] = A(nade.w_enc)[:, 0, :], A(nade.w_dec)[:, :, 0]
    fc_out = rng.standard_normal((N, M * (D + H)))
    fc_out[:, M * H:] -= 1.5
    fc_out = r32(fc_out)
    b_enc, b_dec = g._build_biases(tf.constant(fc_out))
    # targets as RnnMultiNADE.build does it (:118-124): [N, D*M] -> [N, D, M] -> unstack(axis=-1)
    targets = (rng.random((N, D * M)) < 0.08).astype(np.float64)
    tg = tf.unstack(tf.reshape(tf.constant(targets), [-1, D, M]), axis=-1)
    log_prob, cond_prob = [], []
    for m in range(M):
        lp, cp = g._nades[m].log_prob(tg[m], b_enc[m], b_dec[m])
        log_prob.append(lp)
        cond_prob.append(cp)
    outputs = [tf.to_float(tf.greater_equal(cp, .5)) for cp in cond_prob]
    metrics, _, _ = g.build_metrics(tg, outputs, cond_prob, log_prob)
    u = rng.random((M, N, D))
    tfp.push_uniforms([u[m][:, i:i + 1] for m in range(M) for i in range(D)])
    state = RnnEstimatorStateTuple(b_enc, b_dec, None)
    sample, _ = g.sample_single(None, state)
    assert tfp.pending() == 0
    put('multinade', fc_out=fc_out, targets=targets, nll=np.stack([A(l) for l in log_prob], 1),
        cond_p=np.stack([A(c) for c in cond_prob], 0), b_enc=np.stack([A(b) for b in b_enc]),
        b_dec=np.stack([A(b) for b in b_dec]), batch_loss=metrics['batch/loss'],
        log_likelihood=metrics['log_likelihood'], u=u, sample=sample, **w)


def _reference_model(mode, x, lengths, is_train, H=128, R=(48, 32), feedback=(40, 24), keep_prob=1.0):
    """MultINN(config, params, mode) from the reference's own models/multinn/*.py, built twice: the first build creates
    the variables (lazily, inside the cells / Dense layers), which are then rounded to float32-representable values in
    place; the second build recomputes the whole graph with them (the stub executes eagerly)."""
    import yaml
    from models.multinn.multinn import MultINN
    with open(os.path.join(REF, 'configs', 'default_config.yaml')) as f:
        config = yaml.safe_load(f)
    with open(os.path.join(REF, 'configs', 'default_params.yaml')) as f:
        params = yaml.safe_load(f)
    config['training']['num_pixels'] = 1                       # D = 84, the shape BASELINE.json's configs use
    params['generator'].update(num_hidden=H, num_hidden_rnn=list(R), feedback=list(feedback))
    params['keep_prob'] = keep_prob
    tf.reset_default_graph()
    del tf._variables[:]
    tf.feed(x=x, lengths=lengths, is_train=False)
    model = MultINN(config, params, mode=mode, name='multinn')
    for v in tf._variables:
        v[...] = r32(A(v))
    return model


def _rebuild(model, x, lengths, is_train, seed):
    """Re-runs the reference's build() on new placeholder values (the encoders / generators keep their variables)."""
    core = model._model
    core._x, core._lengths = tf.constant(x), tf.constant(A(lengths))
    core._is_train[...] = is_train                             # in place: tf.cond(is_train, ...) holds this placeholder
    for e in core.encoders:                                    # PassEncoder keeps references to the inputs it was built on
        e._is_built = False
    tf.seed_dropout(seed)
    n_vars = len(tf._variables)
    core.build(mode='eval')
    assert len(tf._variables) == n_vars, 'the rebuild must reuse the variables'
    gens = core.generators
    out = dict(loss=A([g.metrics['batch/loss'] for g in gens]),
               predictions=np.stack([A(p) for g in gens for p in (g.forward() if isinstance(g.forward(), list) else [g.forward()])]),
               global_loss=core.metrics['loss'], global_accuracy=core.metrics['accuracy'],
               global_precision=core.metrics['precision'], global_recall=core.metrics['recall'])
    for i, u in enumerate(tf.dropout_log):
        out[f'drop{i}'] = u
    out['n_drop'] = len(tf.dropout_log)
    return out


def mode_cases():
    """The MODE classes (models/multinn/multinn_{composer,jamming,feedback,feedback_rnn}.py + core/) built by the
    reference's own code: input padding and per-track unstack (core/multi_encoder_nn.py:66-76), stack / shift
    (multinn_composer.py:73-87, multinn_jamming.py:60-68, multinn_feedback.py:67-94), the LSTM stack under
    dynamic_decode / dynamic_rnn, Dense, bias split, NADE loop, flatten with ragged lengths, per-track and global
    metrics, dropout placement (is_train=True) and generate() with injected sampler uniforms."""
    modes = {}
    rng = _R32(77)
    B, T, D, M, S = 3, 5, 84, 5, 3
    x = (rng.random((B, T, D, M)) < 0.12).astype(np.float64)
    full, ragged = A([T] * B), A([T, 2, 4])
    for mode in ('composer', 'jamming', 'feedback', 'feedback-rnn'):
        model = _reference_model(mode, x, full, False, keep_prob=0.8)
        key = mode.replace('-', '_')
        for v in tf._variables:
            modes[f'{key}/var/{v.name[:-2]}'] = A(v)
        for case, lengths, is_train in (('eval', full, False), ('ragged', ragged, False), ('train', ragged, True)):
            for k, val in _rebuild(model, x, lengths, is_train, seed=5).items():
                modes[f'{key}/{case}/{k}'] = A(val)
        # generation: intro = the first 3 steps (multinn_composer.py:114-151 and siblings), uniforms u[S, M, D, B]
        _rebuild(model, x[:, :3], A([3] * B), False, seed=5)
        u = rng.random((S, M, D, B))
        # draw order of the reference: Jamming runs each track's whole generate() in turn (multinn_jamming.py:117-125),
        # the other modes sample all tracks inside one step; u[s, m, i] is always (step, track, dimension)
        order = [(s, m) for m in range(M) for s in range(S)] if mode == 'jamming' else \
                [(s, m) for s in range(S) for m in range(M)]
        tfp.push_uniforms([u[s, m, i][:, None] for s, m in order for i in range(D)])
        music = model.generate(S)
        assert tfp.pending() == 0
        modes[f'{key}/generate/u'] = u
        modes[f'{key}/generate/music'] = A(music)
    modes['x'] = x
    modes['ragged'] = ragged
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_modes.npz')
    modes = {k: (v.astype(np.float32) if '/var/' in k or '/drop' in k or k.endswith('/u') else
                 v.astype(np.uint8) if k.endswith(('/predictions', '/music')) or k == 'x' else v) for k, v in modes.items()}
    np.savez_compressed(path, **modes)
    print(f'{path}: {len(modes)} arrays, {os.path.getsize(path) / 1024:.0f} KiB')


def dbn_mode_cases():
    """Composer / Jamming / Joint over DBN encoders with RNN-NADE generators, from the reference's own classes
    (encoders/dbn_encoder.py, common/dbn.py, multinn_joint.py + the files of mode_cases): the per-track (Joint: stacked)
    zero-padded inputs are encoded to SAMPLED codes, the generators model the codes, generated codes are decoded.
    Every Bernoulli draw comes from the stub's seeded fallback and is logged IN CALL ORDER (`draw{i}`), which is how the
    draw order of the reference's graph is discovered rather than assumed:
      build():    per encoder [encode layer 0, layer 1, reconstruct layer 1, layer 0]  (encoder.build, rows b*(T+1) + t),
                  then per encoder the decode of the generator's predictions [layer 1, layer 0]  (rows n = b*T + t)
      generate(): S x M x E sampler draws [B,1] (Jamming: track-major), then per encoder the decode [layer 1, layer 0]
                  (rows b*S + s); the intro encodings are the ones build() sampled.
    The RNN-RBM generator is not covered: the reference's RnnRBM cannot be constructed (`_init_estimator` reads `self.k`
    before `self._k` is assigned, rnn_rbm.py:41-52) and its sampling call passes k=None into tf.constant (rnn_rbm.py:295 ->
    common/rbm.py:223); Joint build() itself only finishes because the stub's merged summary accepts item assignment
    (multinn_core.py:242 on multinn_joint.py:177-186's merged summary raises TypeError under real TF)."""
    import yaml
    from models.multinn.multinn import MultINN
    with open(os.path.join(REF, 'configs', 'default_config.yaml')) as f:
        config = yaml.safe_load(f)
    with open(os.path.join(REF, 'configs', 'default_params.yaml')) as f:
        params = yaml.safe_load(f)
    config['training']['num_pixels'] = 1
    params['generator'].update(type='NADE', num_hidden=128, num_hidden_rnn=[48, 32], feedback=[40, 24])
    params['encoder'].update(type='DBN', num_hidden=[96, 84])
    params['keep_prob'] = 1.0
    rng = _R32(91)
    B, T, D, M, S = 3, 4, 84, 5, 2
    x = (rng.random((B, T, D, M)) < 0.12).astype(np.float64)
    res = {'x': x.astype(np.uint8)}
    for mode in ('composer', 'jamming', 'joint', 'feedback', 'feedback-rnn'):     # feedback-rnn + DBN + NADE = config C4
        key = mode.replace('-', '_')
        tf.reset_default_graph()
        del tf._variables[:]
        tf.feed(x=x, lengths=A([T] * B), is_train=False)
        tfp.auto_uniforms(7)
        model = MultINN(config, params, mode=mode, name='multinn')     # first build: creates the lazy variables
        for v in tf._variables:
            v[...] = r32(A(v))
        core = model._model
        for e in core.encoders:
            e._is_built = False
        tfp.auto_uniforms(11)
        n_vars = len(tf._variables)
        core.build(mode='eval')
        assert len(tf._variables) == n_vars
        for v in tf._variables:
            res[f'{key}/var/{v.name[:-2]}'] = A(v).astype(np.float32)
        res[f'{key}/eval/loss'] = A([g.metrics['batch/loss'] for g in core.generators])
        res[f'{key}/eval/global_loss'] = A(core.metrics['batch/loss'])
        for i, u in enumerate(tfp.auto_log):
            res[f'{key}/eval/draw{i}'] = u.astype(np.float32)
        res[f'{key}/eval/n_draw'] = A(len(tfp.auto_log))
        n0 = len(tfp.auto_log)
        music = model.generate(S)
        draws = tfp.auto_log[n0:]
        E = 84
        n_s = S * M * E if mode != 'joint' else S * E
        sampler = [u for u in draws if u.shape == (B, 1)]          # Jamming interleaves: per track [S*E sampler, 2 decode]
        decode = [u for u in draws if u.shape != (B, 1)]
        assert len(sampler) == n_s and len(decode) == 2 * len(core.encoders)
        res[f'{key}/generate/sampler_draws'] = np.concatenate(sampler, axis=1).T.astype(np.float32)   # [n_s, B] in call order
        for i, u in enumerate(decode):
            res[f'{key}/generate/decode_draw{i}'] = u.astype(np.float32)
        res[f'{key}/generate/music'] = A(music).astype(np.uint8)
    tfp.auto_uniforms(None)
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_dbn_modes.npz')
    np.savez_compressed(path, **res)
    print(f'{path}: {len(res)} arrays, {os.path.getsize(path) / 1024:.0f} KiB')


def joint_rbm_case():
    """Joint mode with the RNN-RBM generator (BASELINE config C3: multinn_joint.py + generators/rnn_rbm.py + common/rbm.py +
    encoders/dbn_encoder.py). The reference cannot run this at HEAD (`RnnRBM.__init__` reads `self.k`
    before `self._k` exists; `sample_single` passes k=None into `tf.constant`), so TWO stand-ins for the evident intent are
    applied from outside, the reference files stay untouched: the class attribute `RnnRBM._k` (what `self._k = k` would have
    set before `_init_estimator` runs) and a wrapper that gives `RBM.sample` the RBM's own `_k` when called without `k` (its
    docstring: "or None if the internal value of k should be used"). Everything else -- stacking and padding of the
    inputs, DBN encoding to sampled codes, `dynamic_rnn` over the LSTM stack, `bh_t = bh + o.Wuh`, `bv_t = bv + o.Wuv`,
    the Gibbs chain from the INPUT frame, the free-energy cost, generation and decoding -- is the reference's code.
    Draws are logged in call order: build(): encode [2], reconstruct [2], chain k x [h, v], decode of the predictions [2];
    generate(): per step k x [h, v], then the decode [2]."""
    import functools
    import yaml
    from models.common.rbm import RBM as RefRBM
    from models.generators.rnn_rbm import RnnRBM
    from models.multinn.multinn import MultINN
    K = 4
    had_k = '_k' in RnnRBM.__dict__
    orig_sample = RefRBM.sample

    @functools.wraps(orig_sample)
    def sample_with_own_k(self, v, bh=None, bv=None, k=None):
        return orig_sample(self, v, bh, bv, self._k if k is None else k)

    RnnRBM._k = K
    RefRBM.sample = sample_with_own_k
    try:
        with open(os.path.join(REF, 'configs', 'default_config.yaml')) as f:
            config = yaml.safe_load(f)
        with open(os.path.join(REF, 'configs', 'default_params.yaml')) as f:
            params = yaml.safe_load(f)
        config['training']['num_pixels'] = 1
        params['generator'].update(type='RBM', num_hidden=64, num_hidden_rnn=[48, 32])
        params['encoder'].update(type='DBN', num_hidden=[96, 84])
        params['keep_prob'] = 1.0
        rng = _R32(97)
        B, T, D, M, S = 3, 4, 84, 5, 2
        x = (rng.random((B, T, D, M)) < 0.12).astype(np.float64)
        res = {'x': x.astype(np.uint8), 'k': A(K)}
        tf.reset_default_graph()
        del tf._variables[:]
        tf.feed(x=x, lengths=A([T] * B), is_train=False)
        tfp.auto_uniforms(13)
        model = MultINN(config, params, mode='joint', name='multinn')
        rs = np.random.default_rng(5)
        for v in tf._variables:
            if v.name.endswith(('/bh:0', '/bv:0')) and 'rnn-rbm' in v.name:      # non-zero RBM biases (zeros at init)
                v[...] = rs.standard_normal(v.shape) * 0.1
            v[...] = r32(A(v))
        core = model._model
        for e in core.encoders:
            e._is_built = False
        tfp.auto_uniforms(17)
        n_vars = len(tf._variables)
        core.build(mode='eval')
        assert len(tf._variables) == n_vars
        for v in tf._variables:
            res[f'var/{v.name[:-2]}'] = A(v).astype(np.float32)
        gen = core.generators[0]
        res['eval/loss'] = A(gen.metrics['batch/loss'])
        res['eval/log_likelihood'] = A(gen.metrics['log_likelihood'])
        res['eval/sample'] = A(gen.forward()).astype(np.uint8)
        assert len(tfp.auto_log) == 4 + 2 * K + 2
        for i, u in enumerate(tfp.auto_log):
            res[f'eval/draw{i}'] = u.astype(np.float32)
        n0 = len(tfp.auto_log)
        music = model.generate(S)
        draws = tfp.auto_log[n0:]
        assert len(draws) == S * 2 * K + 2
        for i, u in enumerate(draws):
            res[f'generate/draw{i}'] = u.astype(np.float32)
        res['generate/music'] = A(music).astype(np.uint8)
    finally:
        RefRBM.sample = orig_sample
        if not had_k:
            del RnnRBM._k
        tfp.auto_uniforms(None)
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_joint_rbm.npz')
    np.savez_compressed(path, **res)
    print(f'{path}: {len(res)} arrays, {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    tf.set_random_seed(20261018)
    nade_cases()
    rbm_cases()
    dbn_cases()
    sequence_cases()
    multinade_cases()
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_primitives.npz')
    np.savez_compressed(path, **out)
    print(f'{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB')
    mode_cases()
    dbn_mode_cases()
    joint_rbm_case()
