"""Generates tests/golden/*.npz: small seeded input/output vectors of the hot path, computed with the NumPy fp64
restatement (oracle/np_oracle.py).

The reference itself cannot run here (tensorflow==1.13.1 has no Python 3.12 wheel, no network) and ships no
golden vectors, so these fixtures pin the ORACLE against regressions and give the CPU and GPU suites one shared
set of known answers; they are not outputs of TF ("parity unpinned").

  python tools/make_golden.py        # rewrites tests/golden/
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import np_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def flat_params(p):
    d = {}
    for l, (k, b) in enumerate(p['lstm']):
        d[f'lstm{l}_kernel'], d[f'lstm{l}_bias'] = k, b
    d['dense_kernel'], d['dense_bias'] = p['dense']
    for m, (we, wd) in enumerate(p['nade']):
        d[f'nade{m}_w_enc'], d[f'nade{m}_w_dec'] = we, wd
    return d


def nade_case():
    rng = np.random.default_rng(101)
    N, D, H = 48, 84, 128
    x = (rng.random((N, D)) < 0.08).astype(np.float64)
    x[0] = 0.0          # empty row (one segment)
    x[1] = 1.0          # full row (D segments)
    x[2, :] = 0.0
    x[2, D - 1] = 1.0   # only the last bit set (opens no segment any dim reads)
    b_enc = rng.standard_normal((N, H))
    b_dec = rng.standard_normal((N, D)) - 2.0
    w_enc = rng.standard_normal((D, H)) / np.sqrt(D)
    w_dec = rng.standard_normal((D, H)) / np.sqrt(D)
    nll, p = O.nade_log_prob(x, b_enc, b_dec, w_enc, w_dec)
    u = rng.random((N, D))
    v, nll_s = O.nade_sample(b_enc, b_dec, w_enc, w_dec, u)
    vt, _ = O.nade_sample(b_enc, b_dec, w_enc, w_dec, None)
    np.savez_compressed(os.path.join(OUT, 'nade.npz'), x=x, b_enc=b_enc, b_dec=b_dec, w_enc=w_enc, w_dec=w_dec,
                        nll=nll, cond_p=p, u=u, sample=v, sample_nll=nll_s, sample_threshold=vt)


def lstm_case():
    rng = np.random.default_rng(102)
    T, B, I, R = 6, 5, 20, (16, 8)
    layers = []
    i = I
    for r in R:
        layers.append((rng.standard_normal((i + r, 4 * r)) * 0.3, rng.standard_normal(4 * r) * 0.1))
        i = r
    x = (rng.random((B, T, I)) < 0.2).astype(np.float64)
    u = [rng.random((T, B, r)) for r in R]
    outs1, st1 = O.rnn_scan(x, layers)
    outs2, st2 = O.rnn_scan(x, layers, keep=0.8, u=u)
    d = dict(x=x, outs_keep1=outs1, outs_keep08=outs2, u0=u[0], u1=u[1])
    for l, (k, b) in enumerate(layers):
        d[f'kernel{l}'], d[f'bias{l}'] = k, b
        d[f'c{l}_keep1'], d[f'h{l}_keep1'] = st1[l]
    np.savez_compressed(os.path.join(OUT, 'lstm.npz'), **d)


def composer_case():
    B, T, D, M, H, R = 4, 6, 84, 5, 128, (64, 32)
    p = O.init_composer_params(D, M, H, R, seed=7)
    x = O.synthetic_pianoroll(B, T, D, M, density=0.06, seed=11).astype(np.float64)
    ref = O.composer_forward(x, O.cast_params(p, np.float64))
    u = np.random.default_rng(12).random((5, M, B, D))
    gen = O.composer_generate(x, O.cast_params(p, np.float64), 5, u)
    d = flat_params(p)
    d.update(x=x, nll=ref['nll'], loss=np.float64(ref['loss']), u=u, generated=gen)
    np.savez_compressed(os.path.join(OUT, 'composer.npz'), **d)


def composer_lengths_case():
    """Variable sequence lengths (utils/sequences.py:6-37) on the composer case's inputs and weights: kept rows only."""
    g = np.load(os.path.join(OUT, 'composer.npz'))
    B, T = g['x'].shape[:2]
    lengths = np.array([T, 2, T - 1, 4][:B], dtype=np.int64)
    p = dict(lstm=[(g[f'lstm{l}_kernel'], g[f'lstm{l}_bias']) for l in range(2)], dense=(g['dense_kernel'], g['dense_bias']),
             nade=[(g[f'nade{m}_w_enc'], g[f'nade{m}_w_dec']) for m in range(5)])
    ref = O.composer_forward(g['x'], O.cast_params(p, np.float64), lengths=lengths)
    np.savez_compressed(os.path.join(OUT, 'composer_lengths.npz'), lengths=lengths, nll=ref['nll'],
                        loss=np.float64(ref['loss']), rows=O.flatten_valid_rows(lengths, T))


def rbm_case():
    rng = np.random.default_rng(103)
    N, D, H, k = 24, 84, 32, 3
    W = rng.standard_normal((D, H)) * 0.2
    bh, bv = rng.standard_normal((1, H)) * 0.1, rng.standard_normal((1, D)) * 0.1 - 1.0
    v = (rng.random((N, D)) < 0.1).astype(np.float64)
    uh, uv = rng.random((k, N, H)), rng.random((k, N, D))
    pv, vk = O.rbm_gibbs(v, W, bh, bv, k, uh, uv)
    F = O.rbm_free_energy(v, W, bh, bv)
    cost = O.rbm_free_energy_cost_mean(v, vk, W, bh, bv)
    np.savez_compressed(os.path.join(OUT, 'rbm.npz'), v=v, W=W, bh=bh, bv=bv, uh=uh, uv=uv, p_v=pv, v_k=vk,
                        free_energy=F, cost=np.float64(cost), k=np.int64(k))


def optim_case():
    rng = np.random.default_rng(104)
    n = 1000
    p, g = rng.standard_normal(n), rng.standard_normal(n) * 3.0
    m, v = np.zeros(n), np.zeros(n)
    (gc,), gn = O.clip_by_global_norm([g], 5.0)
    outs = {}
    pp, mm, vv = p.copy(), m, v
    for t in (1, 2, 3):
        pp, mm, vv = O.tf_adam_step(pp, gc, mm, vv, t)
        outs[f'p{t}'] = pp
    np.savez_compressed(os.path.join(OUT, 'optim.npz'), p=p, g=g, g_clipped=gc, global_norm=np.float64(gn), **outs)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    nade_case()
    lstm_case()
    composer_case()
    composer_lengths_case()
    rbm_case()
    optim_case()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
