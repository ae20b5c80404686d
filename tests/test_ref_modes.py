"""Parity of the MODE-level glue against fixtures produced by the REFERENCE'S OWN MODE CLASSES:
tests/golden/ref_modes.npz is written by tools/make_golden_ref.py::mode_cases, which builds
`MultINN(config, params, mode)` from /root/reference/multinn/models/multinn/{multinn,multinn_composer,multinn_jamming,
multinn_feedback,multinn_feedback_rnn}.py, core/*.py, generators/rnn_{estimator,nade,multinade}.py, common/{rnn,dnn,
nade}.py and encoders/pass_encoder.py -- imported unmodified -- on the NumPy `tensorflow` stand-in of tests/tf_stub.
Zero-padding and track unstacking, the stack / input-target shift, dynamic_decode / dynamic_rnn loops over the LSTM
stack, the Dense bias split, the per-track NADE loop, flattening with ragged lengths, dropout placement, the loss means
and the generate() recurrences are therefore the reference's code; only per-op semantics are NumPy's (float64).

CPU tests: the oracle's mode-level functions reproduce the fixtures (losses to 1e-10, thresholded predictions and
sampled music exactly). GPU tests: the CUDA path, with the fixture's variables loaded through the TF-name importer,
reproduces them (fp32: losses 1e-4 relative; generated music bit-exact with the same uniforms).
Nothing here reads /root/reference at run time."""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import torch_ref as R

Z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_modes.npz'))
TRACKS = ['Drums', 'Piano', 'Guitar', 'Bass', 'Strings']
MODES = ['composer', 'jamming', 'feedback', 'feedback_rnn']
X = Z['x'].astype(np.float64)
B, T, D, M = X.shape
KEEP = 0.8
H, RNN, FEEDBACK = 128, (48, 32), (40, 24)


def variables(mode):
    p = f'{mode}/var/'
    return {k[len(p):]: Z[k].astype(np.float64) for k in Z.files if k.startswith(p)}


def lstm_layers(v, scope, n):
    return [(v[f'{scope}/multi_rnn_cell/cell_{i}/cudnn_compatible_lstm_cell/kernel'],
             v[f'{scope}/multi_rnn_cell/cell_{i}/cudnn_compatible_lstm_cell/bias']) for i in range(n)]


def oracle_params(mode):
    """The fixture's TF variables as the oracle's parameter structures."""
    v = variables(mode)
    if mode == 'composer':
        g = 'multinn/rnn-multinade'
        return dict(lstm=lstm_layers(v, g, 2), dense=(v[f'{g}/dense/kernel'], v[f'{g}/dense/bias']),
                    nade=[(v[f'{g}/all/{t}/nade/w_enc'][:, 0, :], v[f'{g}/all/{t}/nade/w_dec'][:, :, 0]) for t in TRACKS])
    gens = []
    for t in TRACKS:
        g = f'multinn/rnn-nade/{t}'
        gens.append(dict(lstm=lstm_layers(v, g, 2), dense=(v[f'{g}/dense/kernel'], v[f'{g}/dense/bias']),
                         nade=(v[f'{g}/nade/w_enc'][:, 0, :], v[f'{g}/nade/w_dec'][:, :, 0])))
    if mode == 'jamming':
        return gens
    if mode == 'feedback':
        fb = [(v[f'multinn/feedback/dense{s}/kernel'], v[f'multinn/feedback/dense{s}/bias']) for s in ('', '_1')]
    else:
        fb = lstm_layers(v, 'multinn/feedback', 2)
    return gens, fb


def dropout_uniforms(mode, case):
    """The logged DropoutWrapper draws in call order -> (u_fb, u_drop): u_fb = per-layer [T+1,B,F_l] of the feedback RNN
    (built first, multinn_feedback.py:75-80, over the zero-padded T+1 steps), u_drop[g] = per-layer [T,B,R_l] of
    generator g; inside one dynamic_decode / dynamic_rnn the order is step-major, layer-minor."""
    n = int(Z[f'{mode}/{case}/n_drop'])
    if n == 0:
        return None, None
    log = [Z[f'{mode}/{case}/drop{i}'].astype(np.float64) for i in range(n)]
    L = 2
    u_fb = None
    if mode == 'feedback_rnn':
        head, log = log[:(T + 1) * L], log[(T + 1) * L:]
        u_fb = [np.stack(head[l::L]) for l in range(L)]
    G = 1 if mode == 'composer' else M
    assert len(log) == G * T * L
    u_drop = [[np.stack(log[g * T * L:(g + 1) * T * L][l::L]) for l in range(L)] for g in range(G)]
    return u_fb, u_drop


def tt(a):
    if isinstance(a, (list, tuple)):
        return type(a)(tt(b) for b in a)
    if isinstance(a, dict):
        return {k: tt(b) for k, b in a.items()}
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64)


def padded_tracks(x):
    """core/multi_encoder_nn.py:66-76: one zero step in front, per track [B,T+1,D]."""
    return [np.concatenate([np.zeros((x.shape[0], 1, x.shape[2])), x[..., m]], axis=1) for m in range(x.shape[3])]


CASES = {'eval': (None, 1.0), 'ragged': (Z['ragged'], 1.0), 'train': (Z['ragged'], KEEP)}


def oracle_losses(mode, case):
    lengths, keep = CASES[case]
    u_fb, u_drop = dropout_uniforms(mode, case)
    if mode == 'composer':
        r = O.composer_forward(X, oracle_params(mode), keep=keep, u_drop=None if u_drop is None else u_drop[0],
                               lengths=lengths)
        return np.array([r['loss']]), (r['cond_p'] >= 0.5)
    if mode == 'jamming':
        _, nll = R.jamming_loss(tt(X), tt(oracle_params(mode)), keep, tt(u_drop) if u_drop else None, lengths)
    else:
        gens, fb = oracle_params(mode)
        _, nll = R.feedback_loss(tt(padded_tracks(X)), tt(gens), tt(fb), 'dense' if mode == 'feedback' else 'rnn', keep,
                                 tt(u_drop) if u_drop else None, tt(u_fb) if u_fb else None, lengths)
    return nll.mean(0).numpy(), None


# ----------------------------------------------------------------------------- CPU: oracle == reference mode classes
@pytest.mark.parametrize('case', list(CASES))
@pytest.mark.parametrize('mode', MODES)
def test_oracle_mode_loss_matches_reference_code(mode, case):
    loss, pred = oracle_losses(mode, case)
    np.testing.assert_allclose(loss, Z[f'{mode}/{case}/loss'], rtol=1e-10)
    if pred is not None:
        np.testing.assert_array_equal(pred, Z[f'{mode}/{case}/predictions'].astype(bool))


def test_train_case_consumed_dropout():
    """The `train` fixtures really ran with dropout: T*L draws per LSTM stack (+ (T+1)*L for the feedback RNN), and the
    losses differ from the is_train=False ones."""
    for mode, n in (('composer', 10), ('jamming', 50), ('feedback', 50), ('feedback_rnn', 62)):
        assert int(Z[f'{mode}/train/n_drop']) == n
        assert np.all(Z[f'{mode}/train/loss'] != Z[f'{mode}/ragged/loss'])


@pytest.mark.parametrize('mode', MODES)
def test_oracle_generate_matches_reference_code(mode):
    u = Z[f'{mode}/generate/u'].astype(np.float64).transpose(0, 1, 3, 2)          # [S,M,D,B] -> [S,M,B,D]
    S = u.shape[0]
    intro = X[:, :3]
    if mode == 'composer':
        music = O.composer_generate(intro, oracle_params(mode), S, u)
    elif mode == 'jamming':
        music = np.stack([O.composer_generate(intro[..., m:m + 1], _as_multi(p), S, u[:, m:m + 1])[..., 0]
                          for m, p in enumerate(oracle_params(mode))], axis=-1)
    else:
        gens, fb = oracle_params(mode)
        music = O.feedback_generate(padded_tracks(intro), gens, fb, 'dense' if mode == 'feedback' else 'rnn', S, u)
    np.testing.assert_array_equal(music, Z[f'{mode}/generate/music'])


def _as_multi(p):
    """A single-track RNN-NADE as a one-track MultiNADE for composer_generate ([H | D] split is the same for M = 1)."""
    return dict(lstm=p['lstm'], dense=p['dense'], nade=[p['nade']])


def test_global_metrics_of_thresholded_predictions():
    """multi_encoder_nn.py:117-152 + pass_encoder.py:77-92: the `global` metrics are per-track tf.losses.log_loss
    (eps 1e-7) sums between the targets and the THRESHOLDED predictions, averaged over tracks."""
    from multinn_b200.metrics.statistical import global_reconstruction_metrics
    for case in CASES:
        lengths = CASES[case][0]
        rows = np.arange(B * T) if lengths is None else O.flatten_valid_rows(lengths, T)
        tgt = X.transpose(3, 0, 1, 2).reshape(M, B * T, D)[:, rows]                # track-major, rows n = b*T + t
        got = global_reconstruction_metrics(tgt, Z[f'composer/{case}/predictions'].astype(np.float64))
        for k in ('loss', 'accuracy', 'precision', 'recall'):
            np.testing.assert_allclose(got[k], Z[f'composer/{case}/global_{k}'], rtol=1e-10, err_msg=f'{case} {k}')


# ----------------------------------------------------------------------------- GPU: CUDA path == reference mode classes
def _model(mode, device='cuda'):
    """The package's model of `mode` with the fixture's variables loaded BY THEIR TF NAMES (utils/tf_import.py): the names
    come from the reference's own variable_scope / name_scope calls as the stub scopes them."""
    from multinn_b200.multinn import MultINN, default_config, default_params
    from multinn_b200.utils.tf_import import load_tf_variables
    name = mode.replace('_', '-')
    kw = {} if mode in ('composer', 'jamming') else {'feedback': list(FEEDBACK)}
    m = MultINN(default_config(), default_params(mode=name, num_hidden=H, num_hidden_rnn=RNN, keep_prob=KEEP, **kw), name,
                device=device)
    load_tf_variables(m, variables(mode), strict=True)
    return m


@pytest.mark.parametrize('mode', MODES)
def test_reference_named_variables_import_strictly(mode):
    """Every variable the reference's mode class created is consumed, every parameter of the package's model is filled,
    and spot-checked tensors land where the oracle's parameter structures say they belong."""
    m = _model(mode, device='cpu')
    sd = {k: v.numpy() for k, v in m.arena.state_dict().items()}
    assert sum(v.size for v in sd.values()) == sum(v.size for v in variables(mode).values())
    p = oracle_params(mode)
    if mode == 'composer':
        np.testing.assert_array_equal(sd['generator/rnn/cell_1/kernel'], p['lstm'][1][0].astype(np.float32))
        np.testing.assert_array_equal(sd['generator/nade/w_dec'][3], p['nade'][3][1].astype(np.float32))
    else:
        gens = p if mode == 'jamming' else p[0]
        np.testing.assert_array_equal(sd['generator/Guitar/rnn/cell_0/kernel'], gens[2]['lstm'][0][0].astype(np.float32))
        np.testing.assert_array_equal(sd['generator/Bass/dense/kernel'], gens[3]['dense'][0].astype(np.float32))


def _drop_for_model(mode, case):
    """The logged draws in the layout the package's step() takes: u_drop[g][l] = [T,B,R_l] (CUDA tensors)."""
    u_fb, u_drop = dropout_uniforms(mode, case)
    f = lambda a: torch.tensor(a, dtype=torch.float32).cuda()
    return (None if u_fb is None else [f(a) for a in u_fb],
            None if u_drop is None else [[f(a) for a in g] for g in u_drop])


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['eval', 'ragged'])
@pytest.mark.parametrize('mode', MODES)
def test_cuda_evaluate_matches_reference_code(mode, case):
    m = _model(mode)
    lengths = CASES[case][0]
    x = torch.tensor(X, dtype=torch.float32).cuda()
    out = m.evaluate(x, lengths=None if lengths is None else torch.tensor(lengths))
    per_track = out['nll'].mean(0).cpu().numpy().astype(np.float64)
    ref = Z[f'{mode}/{case}/loss']
    np.testing.assert_allclose(per_track.mean() if mode == 'composer' else per_track, ref if mode != 'composer' else ref[0],
                               rtol=1e-4)
    np.testing.assert_allclose(float(out['batch/loss']), ref.mean(), rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('mode', MODES)
def test_cuda_train_forward_matches_reference_code(mode):
    """is_train=True: ragged lengths and the reference's logged dropout draws; the training loss (mean of the track
    losses) of the CUDA forward + backward pass against the reference's."""
    m = _model(mode)
    core = m._model
    u_fb, u_drop = _drop_for_model(mode, 'train')
    lengths = torch.tensor(Z['ragged'])
    x = core._check_x(torch.tensor(X, dtype=torch.float32).cuda(), lengths)
    kw = {'u_fb': u_fb} if mode == 'feedback_rnn' else {}
    core.arena.grad.zero_()
    loss = core._forward_backward(x, keep=KEEP, u_drop=u_drop[0] if mode == 'composer' else u_drop, seed=0, lengths=lengths,
                                  **kw)
    np.testing.assert_allclose(float(loss), Z[f'{mode}/train/loss'].mean(), rtol=1e-4)
    assert float(core.arena.grad.abs().sum()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize('mode', MODES)
def test_cuda_generate_matches_reference_code(mode):
    m = _model(mode)
    u = torch.tensor(Z[f'{mode}/generate/u'].transpose(0, 1, 3, 2).copy(), dtype=torch.float32).cuda()   # [S,M,B,D]
    x = torch.tensor(X[:, :3].copy(), dtype=torch.float32).cuda()
    music = m.generate(x, u.shape[0], u=u)
    np.testing.assert_array_equal(music.cpu().numpy().astype(np.uint8), Z[f'{mode}/generate/music'])
