"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py: pinned by the reference's own modules run on tests/tf_stub).

Restatement (ii): torch-CPU form with autograd, op for op like the TF1 graph (python-unrolled
84-iteration NADE loop per track as in common/nade.py:199-226, per-step LSTM loop as in
dynamic_decode, generators/rnn_nade.py:204-218, autograd backward, global-norm clip 5 and
TF-style Adam as in utils/training.py:163-175 / train.py:64). It is the gradient oracle for the
CUDA path and the timed "TF1-graph restatement" CPU baseline of bench.py (TF 1.13.1 is not
installable here).
"""
import math

import numpy as np
import torch

EPS = 1e-6


def safe_log(t):
    """utils/auxiliary.py:9-11."""
    return torch.log(EPS + t)


def nade_log_prob(x, b_enc, b_dec, w_enc, w_dec):
    """common/nade.py:155-229, unrolled loop. Returns (nll[N], cond_p[N,D])."""
    N, D = x.shape
    a = b_enc
    log_p = torch.zeros(N, 1, dtype=x.dtype)
    cond = []
    for i in range(D):
        h = torch.sigmoid(a)
        l_i = b_dec[:, i:i + 1] + h @ w_dec[i][:, None]
        p_i = torch.sigmoid(l_i)
        v_i = x[:, i:i + 1]
        log_p = log_p + (v_i * safe_log(p_i) + (1 - v_i) * safe_log(1 - p_i))
        cond.append(p_i)
        a = a + v_i @ w_enc[i][None, :]
    return -log_p[:, 0], torch.cat(cond, 1)


def lstm_cell(x, c, h, kernel, bias):
    """common/rnn.py:124 (LSTMBlockCell, forget_bias 0, gate order i,j,f,o)."""
    g = torch.cat([x, h], 1) @ kernel + bias
    i, j, f, o = g.chunk(4, 1)
    c2 = torch.tanh(j) * torch.sigmoid(i) + c * torch.sigmoid(f)
    h2 = torch.tanh(c2) * torch.sigmoid(o)
    return c2, h2


def dropout(x, keep, u):
    """tf.nn.dropout (common/rnn.py:117-132)."""
    if u is None or keep >= 1.0:
        return x
    return x / keep * torch.floor(keep + u)


def rnn_scan(inputs, layers, keep=1.0, u=None, state=None):
    """T-loop of MultiRNNCell steps. u: list per layer of [T,B,R]. Returns (outs[B,T,R], state)."""
    B, T, _ = inputs.shape
    if state is None:
        state = [(torch.zeros(B, k.shape[1] // 4, dtype=inputs.dtype),
                  torch.zeros(B, k.shape[1] // 4, dtype=inputs.dtype)) for k, _ in layers]
    outs = []
    for t in range(T):
        inp = inputs[:, t]
        new = []
        for l, (kernel, bias) in enumerate(layers):
            c2, h2 = lstm_cell(inp, state[l][0], state[l][1], kernel, bias)
            new.append((c2, h2))
            inp = dropout(h2, keep, None if u is None else u[l][t])
        state = new
        outs.append(inp)
    return torch.stack(outs, 1), state


def composer_loss(x, params, keep=1.0, u_drop=None, lengths=None):
    """Composer LSTM-MultiNADE training graph. x[B,T,D,M] tensor.
    params: dict(lstm=[(k,b)], dense=(K,b), nade=[(w_enc,w_dec)]). Returns (loss, nll[N,M]).
    lengths: rows t >= lengths[b] are dropped before the means (utils/sequences.py:6-37)."""
    B, T, D, M = x.shape
    stack = x.reshape(B, T, D * M)
    pad = torch.cat([torch.zeros(B, 1, D * M, dtype=x.dtype), stack], 1)
    inp, tgt = pad[:, :-1], pad[:, 1:]
    outs, _ = rnn_scan(inp, params['lstm'], keep, u_drop)
    K, b = params['dense']
    fc = outs.reshape(B * T, -1) @ K + b
    H = params['nade'][0][0].shape[1]
    tgt_flat = tgt.reshape(B * T, D, M)
    nlls = []
    for m in range(M):
        be = fc[:, m * H:(m + 1) * H]
        bd = fc[:, M * H + m * D:M * H + (m + 1) * D]
        nll, _ = nade_log_prob(tgt_flat[:, :, m], be, bd, *params['nade'][m])
        nlls.append(nll)
    if lengths is not None and int(min(lengths)) != T:
        keep_rows = torch.cat([b * T + torch.arange(int(l)) for b, l in enumerate(lengths)])
        nlls = [n[keep_rows] for n in nlls]
    loss = torch.stack([n.mean() for n in nlls]).mean()
    return loss, torch.stack(nlls, 1)


def rnn_nade_loss(inp, tgt, params, keep=1.0, u_drop=None, lengths=None):
    """Single-track RNN-NADE (Jamming generator), generators/rnn_nade.py:279-302; padded rows dropped
    (rnn_nade.py:225, utils/sequences.py:29-37)."""
    B, T, D = tgt.shape
    outs, _ = rnn_scan(inp, params['lstm'], keep, u_drop)
    K, b = params['dense']
    fc = outs.reshape(B * T, -1) @ K + b
    H = params['nade'][0].shape[1]
    nll, _ = nade_log_prob(tgt.reshape(B * T, D), fc[:, :H], fc[:, H:H + D], *params['nade'])
    if lengths is not None and int(min(lengths)) != T:
        nll = nll[torch.cat([b * T + torch.arange(int(l)) for b, l in enumerate(lengths)])]
    return nll.mean(), nll


def jamming_loss(x, params_list, keep=1.0, u_drop=None, lengths=None):
    """Jamming: M independent RNN-NADEs, loss = mean of track losses (multinn_core.py:402-405)."""
    B, T, D, M = x.shape
    losses, nlls = [], []
    for m in range(M):
        xm = x[..., m]
        pad = torch.cat([torch.zeros(B, 1, D, dtype=x.dtype), xm], 1)
        l, n = rnn_nade_loss(pad[:, :-1], pad[:, 1:], params_list[m], keep,
                             None if u_drop is None else u_drop[m], lengths)
        losses.append(l)
        nlls.append(n)
    return torch.stack(losses).mean(), torch.stack(nlls, 1)


def flat_params(params):
    """Flatten a parameter tree into a list of leaf tensors (deterministic order)."""
    out = []
    if isinstance(params, dict):
        for k in params:
            out += flat_params(params[k])
    elif isinstance(params, (list, tuple)):
        for v in params:
            out += flat_params(v)
    else:
        out.append(params)
    return out


def to_torch(params, dtype=torch.float32, requires_grad=False):
    if isinstance(params, dict):
        return {k: to_torch(v, dtype, requires_grad) for k, v in params.items()}
    if isinstance(params, (list, tuple)):
        return type(params)(to_torch(v, dtype, requires_grad) for v in params)
    t = torch.tensor(np.asarray(params), dtype=dtype)
    t.requires_grad_(requires_grad)
    return t


class TFAdam:
    """tf.train.AdamOptimizer(lr, epsilon=1e-4) (train.py:64) after clip_by_global_norm(5.)
    (utils/training.py:166). SURVEY 9.7."""

    def __init__(self, leaves, lr=0.01, b1=0.9, b2=0.999, eps=1e-4, clip=5.0):
        self.leaves = leaves
        self.lr, self.b1, self.b2, self.eps, self.clip = lr, b1, b2, eps, clip
        self.m = [torch.zeros_like(p) for p in leaves]
        self.v = [torch.zeros_like(p) for p in leaves]
        self.t = 0

    @torch.no_grad()
    def step(self, grads):
        gn = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
        scale = self.clip / max(gn, self.clip)
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for p, g, m, v in zip(self.leaves, grads, self.m, self.v):
            g = g * scale
            m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            p.sub_(lr_t * m / (v.sqrt() + self.eps))
        return gn


def composer_train_step(x, params, opt, keep=1.0, u_drop=None):
    """One sess.run of train.py:186-189: fwd + bwd + clip + Adam. Returns (loss, grad_norm)."""
    leaves = opt.leaves
    loss, _ = composer_loss(x, params, keep, u_drop)
    grads = torch.autograd.grad(loss, leaves)
    gn = opt.step(list(grads))
    return float(loss), gn


# ----------------------------------------------------------------------------- RBM / DBN / feedback (torch, autograd)
def rbm_free_energy(v, W, bh, bv):
    """common/rbm.py:256-258 per row (un-broadcast [N] vector, quirk Q4)."""
    return -torch.nn.functional.softplus(v @ W + bh).sum(1) - (v * bv).sum(1)


def rbm_free_energy_cost_mean(v, v_sample, W, bh, bv):
    """common/rbm.py:233-263 + statistical.py:34 (mean over the [N,N] broadcast)."""
    return rbm_free_energy(v, W, bh, bv).mean() - rbm_free_energy(v_sample, W, bh, bv).mean()


def dnn_forward(x, layers):
    """common/dnn.py:97-116: sigmoid Dense layers."""
    for K, b in layers:
        x = torch.sigmoid(x @ K + b)
    return x


def feedback_loss(xe, gen_params, fb_params, kind='dense', keep=1.0, u_drop=None, u_fb=None, lengths=None):
    """multinn_feedback.py:54-97 / multinn_feedback_rnn.py:41-79 training graph on given per-track encodings.
    xe: list of M tensors [B,T+1,E] (zero-padded, already encoded). gen_params: list of rnn-nade param dicts (LSTM input
    = E + F). fb_params: dense -> [(K,b)..]; rnn -> [(kernel,bias)..]. Returns (loss, nll[N,M])."""
    M = len(xe)
    B, T1, E = xe[0].shape
    T = T1 - 1
    stack = torch.stack(xe, dim=3).reshape(B, T1, E * M)
    if kind == 'dense':
        fb = dnn_forward(stack.reshape(B * T1, -1), fb_params).reshape(B, T1, -1)
    else:
        fb, _ = rnn_scan(stack, fb_params, keep, u_fb)
    losses, nlls = [], []
    for m in range(M):
        inp = torch.cat([xe[m], fb], dim=2)[:, :-1]
        tgt = xe[m][:, 1:]
        l, n = rnn_nade_loss(inp, tgt, gen_params[m], keep, None if u_drop is None else u_drop[m],
                             lengths)       # the generators alone see `lengths` (multinn_feedback.py:93-94)
        losses.append(l)
        nlls.append(n)
    return torch.stack(losses).mean(), torch.stack(nlls, 1)
