"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py: pinned by the reference's own modules run on tests/tf_stub).

Restatement (i): NumPy literal-loop form of the reference hot path. Mirrors the TF1 graph
op for op (D-loop for NADE, T-loop for the LSTM, k-loop for the Gibbs chain). Default dtype
float64; pass dtype=np.float32 to mimic the reference's arithmetic width.

Every function cites the reference file:line (relative to /root/reference/multinn/) it follows.
TF 1.13.1 / TFP 0.6.0 library semantics are those listed in SURVEY.md section 9.
"""
import numpy as np

EPS_SAFE_LOG = 1e-6  # utils/auxiliary.py:11


# ----------------------------------------------------------------------------- primitives
def sigmoid(x):
    """tf.sigmoid (common/nade.py:326-328, common/rbm.py:352,371)."""
    return 1.0 / (1.0 + np.exp(-x))


def safe_log(t):
    """utils/auxiliary.py:9-11: log(1e-6 + t)."""
    return np.log(t.dtype.type(EPS_SAFE_LOG) + t)


def bernoulli_sample(p, u):
    """TFP 0.6.0 Bernoulli._sample_n at common/nade.py:283-287, common/rbm.py:386-387:
    sample = cast(u < probs), strict <, u in [0,1)."""
    return (u < p).astype(p.dtype)


# ----------------------------------------------------------------------------- NADE
def nade_log_prob(x, b_enc, b_dec, w_enc, w_dec):
    """common/nade.py:155-229 (loop_body :199-221, _cond_prob :310-329).

    x[N,D] in {0,1}; b_enc[N,H]; b_dec[N,D]; w_enc[D,H] (= w_enc[D,1,H] squeezed);
    w_dec[D,H] (= w_dec[D,H,1] squeezed). Returns (nll[N] positive, cond_p[N,D]).
    """
    N, D = x.shape
    a = b_enc.copy()                                   # :190
    log_p = np.zeros((N,), dtype=x.dtype)              # :191
    cond_p = np.zeros((N, D), dtype=x.dtype)
    for i in range(D):                                 # :225
        h = sigmoid(a)                                 # :326
        l_i = b_dec[:, i] + h @ w_dec[i]               # :327
        p_i = sigmoid(l_i)                             # :328
        v_i = x[:, i]
        log_p = log_p + (v_i * safe_log(p_i) + (1 - v_i) * safe_log(1 - p_i))  # :210-213
        cond_p[:, i] = p_i                             # :216
        a = a + v_i[:, None] * w_enc[i][None, :]       # :219
    return -log_p, cond_p                              # :228-229


def nade_log_prob_triangular(x, b_enc, b_dec, w_enc, w_dec):
    """Second, structurally different statement of nade.py:155-229:
    A[n,i,:] = b_enc[n] + sum_{j<i} x[n,j] w_enc[j] (strictly lower triangular contraction)."""
    N, D = x.shape
    tri = np.tril(np.ones((D, D), dtype=x.dtype), -1)              # [i,j] = 1 if j<i
    A = b_enc[:, None, :] + np.einsum('ij,nj,jh->nih', tri, x, w_enc)
    L = b_dec + np.einsum('nih,ih->ni', sigmoid(A), w_dec)
    P = sigmoid(L)
    ll = (x * safe_log(P) + (1 - x) * safe_log(1 - P)).sum(1)
    return -ll, P


def nade_sample(b_enc, b_dec, w_enc, w_dec, u=None):
    """common/nade.py:231-308. u[N,D] uniforms -> temperature=1. sampling
    (v_i = float(u_i < sigmoid(l_i / 1.)), :283-287); u=None -> temperature=None -> p>=0.5 (:281).
    Returns (v[N,D], nll[N])."""
    N, D = b_dec.shape
    a = b_enc.copy()
    log_p = np.zeros((N,), dtype=b_enc.dtype)
    v = np.zeros((N, D), dtype=b_enc.dtype)
    for i in range(D):
        h = sigmoid(a)
        l_i = b_dec[:, i] + h @ w_dec[i]
        p_i = sigmoid(l_i)
        if u is None:
            v_i = (p_i >= 0.5).astype(b_enc.dtype)
        else:
            v_i = bernoulli_sample(p_i, u[:, i])
        v[:, i] = v_i
        log_p = log_p + (v_i * safe_log(p_i) + (1 - v_i) * safe_log(1 - p_i))  # :293-296
        a = a + v_i[:, None] * w_enc[i][None, :]                               # :299
    return v, -log_p


# ----------------------------------------------------------------------------- LSTM temporal unit
def lstm_cell(x, c, h, kernel, bias):
    """CudnnCompatibleLSTMCell == LSTMBlockCell(forget_bias=0) at common/rnn.py:124.
    kernel[(I+R),4R] rows = [x ; h]; column blocks i, j(cell input), f, o."""
    R = h.shape[1]
    g = np.concatenate([x, h], axis=1) @ kernel + bias
    i, j, f, o = g[:, :R], g[:, R:2 * R], g[:, 2 * R:3 * R], g[:, 3 * R:]
    c_new = np.tanh(j) * sigmoid(i) + c * sigmoid(f)
    h_new = np.tanh(c_new) * sigmoid(o)
    return c_new, h_new


def dropout(x, keep, u):
    """tf.nn.dropout under DropoutWrapper(output_keep_prob) at common/rnn.py:117-132:
    out = x / keep * floor(keep + u)."""
    if u is None or keep >= 1.0:
        return x
    k = x.dtype.type(keep)
    return x / k * np.floor(k + u.astype(x.dtype))


def multi_rnn_step(x, state, layers, keep=1.0, u=None):
    """MultiRNNCell([DropoutWrapper(LSTM)]) one step (common/rnn.py:122-137).
    state = [(c,h)] per layer; layers = [(kernel,bias)]; u = list of [B,R_l] uniforms or None.
    State h is NOT dropped; the layer output is."""
    new_state = []
    inp = x
    for l, (kernel, bias) in enumerate(layers):
        c, h = state[l]
        c2, h2 = lstm_cell(inp, c, h, kernel, bias)
        new_state.append((c2, h2))
        inp = dropout(h2, keep, None if u is None else u[l])
    return inp, new_state


def zero_state(B, layers, dtype):
    """common/rnn.py:155-176 (learn_zero_state=False)."""
    st = []
    for kernel, _ in layers:
        R = kernel.shape[1] // 4
        st.append((np.zeros((B, R), dtype), np.zeros((B, R), dtype)))
    return st


def rnn_scan(inputs, layers, keep=1.0, u=None, state=None):
    """dynamic_decode(TrainingHelper) loop with full lengths (generators/rnn_nade.py:204-218).
    inputs[B,T,I]; u = list per layer of [T,B,R_l] or None. Returns (outputs[B,T,R_top], state)."""
    B, T, _ = inputs.shape
    if state is None:
        state = zero_state(B, layers, inputs.dtype)
    outs = []
    for t in range(T):
        ut = None if u is None else [ul[t] for ul in u]
        o, state = multi_rnn_step(inputs[:, t], state, layers, keep, ut)
        outs.append(o)
    return np.stack(outs, axis=1), state


def dense(x, kernel, bias):
    """tf.layers.Dense without activation (generators/rnn_nade.py:54-57)."""
    return x @ kernel + bias


# ----------------------------------------------------------------------------- RNN-(Multi)NADE
def split_biases_multi(fc_out, M, H, D):
    """generators/rnn_multinade.py:231-256: [M*H | M*D] then M equal chunks each."""
    be = [fc_out[:, m * H:(m + 1) * H] for m in range(M)]
    bd = [fc_out[:, M * H + m * D: M * H + (m + 1) * D] for m in range(M)]
    return be, bd


def composer_inputs_targets(x):
    """core/multi_encoder_nn.py:66-76 + multinn_composer.py:73-87 (PassEncoder = identity).
    x[B,T,D,M] -> inp[B,T,D*M] = [0, x_0..x_{T-2}], tgt[B,T,D*M] = x; feature = d*M + m."""
    B, T, D, M = x.shape
    stack = x.reshape(B, T, D * M)
    pad = np.concatenate([np.zeros((B, 1, D * M), x.dtype), stack], axis=1)
    return pad[:, :-1], pad[:, 1:]


def flatten_valid_rows(lengths, T):
    """utils/sequences.py:29-31: indices of tf.where(tf.sequence_mask(lengths)) into the [B*T] row order n = b*T + t
    (b-major, rows t >= lengths[b] removed)."""
    lengths = np.asarray(lengths)
    return np.concatenate([b * T + np.arange(int(l)) for b, l in enumerate(lengths)]).astype(np.int64)


def composer_forward(x, params, keep=1.0, u_drop=None, lengths=None):
    """Composer LSTM-MultiNADE teacher-forced forward (SURVEY 3.2 stack).
    params: dict(lstm=[(kernel,bias)..], dense=(K,b), nade=[(w_enc,w_dec)..M]).
    Returns dict(nll[N,M], cond_p[M,N,D], loss, fc_out[N,*]); rows n = b*T + t (utils/sequences.py:22-24).
    lengths[B] (max == T): rows past a sequence's length are dropped from every per-row result and from the loss mean
    (dynamic_decode still computes them; flatten_maybe_padded_sequences removes them)."""
    B, T, D, M = x.shape
    inp, tgt = composer_inputs_targets(x)
    outs, _ = rnn_scan(inp, params['lstm'], keep, u_drop)
    K, b = params['dense']
    fc = dense(outs.reshape(B * T, -1), K, b)                       # rnn_nade.py:212,225
    H = params['nade'][0][0].shape[1]
    be, bd = split_biases_multi(fc, M, H, D)
    tgt_flat = tgt.reshape(B * T, D, M)                             # rnn_multinade.py:111-116
    nll = np.zeros((B * T, M), x.dtype)
    cp = np.zeros((M, B * T, D), x.dtype)
    for m in range(M):                                              # rnn_multinade.py:281-288
        nll[:, m], cp[m] = nade_log_prob(tgt_flat[:, :, m], be[m], bd[m], *params['nade'][m])
    if lengths is not None and int(np.min(lengths)) != T:           # sequences.py:33-37
        keep_rows = flatten_valid_rows(lengths, T)
        nll, cp, fc = nll[keep_rows], cp[:, keep_rows], fc[keep_rows]
    loss = np.mean([nll[:, m].mean() for m in range(M)])            # statistical.py:34; rnn_multinade.py:200-203
    return dict(nll=nll, cond_p=cp, loss=loss, fc_out=fc)


def rnn_nade_forward(inp, tgt, params, keep=1.0, u_drop=None):
    """Single-track RNN-NADE (Jamming generator) forward: generators/rnn_nade.py:279-302.
    inp[B,T,I], tgt[B,T,D]; params dict(lstm, dense, nade=(w_enc,w_dec))."""
    B, T, D = tgt.shape
    outs, _ = rnn_scan(inp, params['lstm'], keep, u_drop)
    fc = dense(outs.reshape(B * T, -1), *params['dense'])
    H = params['nade'][0].shape[1]
    b_enc, b_dec = fc[:, :H], fc[:, H:H + D]                        # rnn_nade.py:245
    nll, cp = nade_log_prob(tgt.reshape(B * T, D), b_enc, b_dec, *params['nade'])
    return dict(nll=nll, cond_p=cp, loss=nll.mean(), fc_out=fc)


def composer_generate(x_intro, params, num_steps, u):
    """generators/rnn_estimator.py:271-323 + rnn_multinade.py:292-317 + multinn_composer.py:114-151.
    x_intro[B,Ti,D,M]; u[S,M,B,D] uniforms (None -> threshold sampling). Returns [B,S,D,M].
    The intro scan covers the zero-padded sequence (Ti+1 steps); dropout off."""
    B, Ti, D, M = x_intro.shape
    stack = x_intro.reshape(B, Ti, D * M)
    pad = np.concatenate([np.zeros((B, 1, D * M), x_intro.dtype), stack], axis=1)
    return multinade_generate(pad, params, num_steps, u).reshape(B, num_steps, D, M)


def multinade_generate(intro, params, num_steps, u):
    """RnnEstimator.generate for an RNN-MultiNADE over given (already padded / encoded) intro features
    intro[B,T1,E*M], feature e*M + m (generators/rnn_estimator.py:271-323, rnn_multinade.py:292-317).
    u[S,M,B,E] or None. Returns samples[B,S,E*M]."""
    B = intro.shape[0]
    M = len(params['nade'])
    E, H = params['nade'][0][0].shape
    outs, state = rnn_scan(intro, params['lstm'])
    K, b = params['dense']
    fc = dense(outs[:, -1], K, b)                                   # last_outputs=True, rnn_nade.py:223
    samples = []
    for s in range(num_steps):
        be, bd = split_biases_multi(fc, M, H, E)
        vs = []
        for m in range(M):
            v, _ = nade_sample(be[m], bd[m], *params['nade'][m], u=None if u is None else u[s, m])
            vs.append(v)
        samp = np.stack(vs, axis=2).reshape(B, E * M)               # rnn_multinade.py:314-315
        samples.append(samp)
        o, state = multi_rnn_step(samp, state, params['lstm'])      # rnn_nade.py:267
        fc = dense(o, K, b)
    return np.stack(samples, axis=1)


# ----------------------------------------------------------------------------- RBM / DBN
def rbm_cond_prob_h(v, W, bh):
    """common/rbm.py:337-354."""
    return sigmoid(v @ W + bh)


def rbm_cond_prob_v(h, W, bv):
    """common/rbm.py:356-373."""
    return sigmoid(h @ W.T + bv)


def rbm_forward(v, W, bh, u):
    """common/rbm.py:148-167."""
    p = rbm_cond_prob_h(v, W, bh)
    return p, bernoulli_sample(p, u)


def rbm_reconstruct(h, W, bv, u):
    """common/rbm.py:169-190."""
    p = rbm_cond_prob_v(h, W, bv)
    return p, bernoulli_sample(p, u)


def rbm_gibbs(v, W, bh, bv, k, uh, uv):
    """common/rbm.py:192-231. uh[k,N,H], uv[k,N,D]. k=0 returns (v, v) like the while_loop init (:222-226).
    Returns (p_v of the last step, v_k)."""
    p_v, vk = v, v
    for s in range(k):
        _, hk = rbm_forward(vk, W, bh, uh[s])
        p_v, vk = rbm_reconstruct(hk, W, bv, uv[s])
    return p_v, vk


def softplus(x):
    return np.logaddexp(0.0, x)


def rbm_free_energy(v, W, bh, bv):
    """F(v) of common/rbm.py:256-258 per row (the un-broadcast, intended [N] vector):
    F(v) = -sum_j log(1+exp((vW+bh)_j)) - v.bv. bh[1,H] or [N,H]; bv[1,D] or [N,D]."""
    return -softplus(v @ W + bh).sum(1) - (v * bv).sum(1)


def rbm_free_energy_cost_mean(v, v_sample, W, bh, bv):
    """common/rbm.py:233-263 + metrics/statistical.py:34: batch/loss = reduce_mean(cost).
    Quirk Q4: the [N]-[N,1] broadcast makes cost [N,N]; its mean equals mean(F(v)) - mean(F(vs))
    when bh,bv are [1,*] (the call site :119 always passes the internal biases)."""
    return rbm_free_energy(v, W, bh, bv).mean() - rbm_free_energy(v_sample, W, bh, bv).mean()


def rbm_cd_update(v, W, bh, bv, k, lr, uh, uv, uh0, uhk):
    """common/rbm.py:299-335. uh[k,N,H], uv[k,N,D] drive the chain; uh0[N,H] samples h~p(h|v);
    uhk[N,H] is drawn for forward(v_sample) (its sample is unused, only p_h_sample is).
    Returns (dW, dbv, dbh) to be assign_add-ed."""
    N = v.shape[0]
    p_vs, vs = rbm_gibbs(v, W, bh, bv, k, uh, uv)
    _, h = rbm_forward(v, W, bh, uh0)
    p_hs, _ = rbm_forward(vs, W, bh, uhk)
    a = v.dtype.type(lr) / v.dtype.type(N)
    dW = a * (v.T @ h - p_vs.T @ p_hs)
    dbv = a * (v - p_vs).sum(0, keepdims=True)
    dbh = a * (h - p_hs).sum(0, keepdims=True)
    return dW, dbv, dbh


def rbm_visible_bias_init(v):
    """common/rbm.py:286-297: bv = log(1e-6 + p/(1-p)), p = mean(v,0)."""
    p = v.mean(0)
    return safe_log(p / (1 - p))[None, :]


def dbn_forward(v, rbms, us):
    """common/dbn.py:136-156; rbms = [(W,bh,bv)], us = list of uniforms per layer."""
    p, h = 0, v
    for (W, bh, _), u in zip(rbms, us):
        p, h = rbm_forward(h, W, bh, u)
    return p, h


def dbn_reconstruct(h, rbms, us):
    """common/dbn.py:158-180; us ordered like the loop (last layer first)."""
    p, v = 0, h
    for (W, _, bv), u in zip(reversed(rbms), us):
        p, v = rbm_reconstruct(v, W, bv, u)
    return p, v


def rnn_rbm_forward(inp, tgt, params, k, uh, uv, keep=1.0, u_drop=None):
    """generators/rnn_rbm.py:94-119 training/eval graph. inp[B,T,I] (I == D for the chain start),
    tgt[B,T,D]. params: lstm, Wuh[R,H], Wuv[R,D], rbm=(W,bh,bv). internal_bias=True (:22).
    Chain starts from the INPUT frame (:112). Loss uses INTERNAL bh,bv (quirk Q3)."""
    B, T, D = tgt.shape
    outs, _ = rnn_scan(inp, params['lstm'], keep, u_drop)
    o = outs.reshape(B * T, -1)
    W, bh, bv = params['rbm']
    bh_t = bh + o @ params['Wuh']                                   # rnn_rbm.py:252-257
    bv_t = bv + o @ params['Wuv']
    p_v, v_s = rbm_gibbs(inp.reshape(B * T, -1), W, bh_t, bv_t, k, uh, uv)
    loss = rbm_free_energy_cost_mean(tgt.reshape(B * T, D), v_s, W, bh, bv)
    return dict(cond_p=p_v, sample=v_s, loss=loss, bh_t=bh_t, bv_t=bv_t)


def rnn_rbm_generate(codes, params, k, num_steps, us):
    """RnnEstimator.generate for an RNN-RBM (generators/rnn_estimator.py:271-323, rnn_rbm.py:261-297): scan the intro
    codes[B,Ti,E] (dropout off) -> last-step biases and RNN state; then `num_steps` times: k-step Gibbs chain STARTED FROM
    THE PREVIOUS FRAME (the last intro frame first, rnn_rbm.py:295) with the state's biases, one LSTM step on the sample,
    new biases. us[s] = (uh[k,B,H], uv[k,B,E]). Returns samples[B,num_steps,E]."""
    W, bh, bv = params['rbm']
    outs, state = rnn_scan(codes, params['lstm'])
    o = outs[:, -1]
    prev = codes[:, -1]
    B, E = prev.shape
    out = np.zeros((B, num_steps, E), codes.dtype)
    for s in range(num_steps):
        bh_t = bh + o @ params['Wuh']                                # rnn_rbm.py:252-257 (internal_bias=True)
        bv_t = bv + o @ params['Wuv']
        _, prev = rbm_gibbs(prev, W, bh_t, bv_t, k, us[s][0], us[s][1])
        out[:, s] = prev
        o, state = multi_rnn_step(prev, state, params['lstm'])
    return out


def joint_generate(x_intro, enc_rbms, gen_params, k, num_steps, u_enc, us, u_dec):
    """multinn_joint.py:188-215: encode the zero-padded stacked intro with the DBN (sampled codes, quirk Q12), generate
    codes with the RNN-RBM, decode them through the DBN (sampled visibles). Rows of the encoder / decoder uniforms are
    TIME-MAJOR (t*B + b) resp. (b*S + s), as the device code lays them out. Returns [B,S,D,M]."""
    B, Ti, D, M = x_intro.shape
    inp, _ = composer_inputs_targets(x_intro)
    pad = np.concatenate([inp, x_intro.reshape(B, Ti, -1)[:, -1:]], axis=1)           # [B,Ti+1,D*M]
    flat_tm = pad.transpose(1, 0, 2).reshape((Ti + 1) * B, -1)
    _, codes = dbn_forward(flat_tm, enc_rbms, u_enc)
    codes = codes.reshape(Ti + 1, B, -1).transpose(1, 0, 2)
    samples_h = rnn_rbm_generate(codes, gen_params, k, num_steps, us)                 # [B,S,E]
    _, v = dbn_reconstruct(samples_h.reshape(B * num_steps, -1), enc_rbms, u_dec)
    return v.reshape(B, num_steps, D, M)


# ----------------------------------------------------------------------------- optimiser
def clip_by_global_norm(grads, clip=5.0):
    """tf.clip_by_global_norm at utils/training.py:166: g * clip / max(gn, clip)."""
    gn = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))
    scale = clip / max(gn, clip)
    return [g * g.dtype.type(scale) for g in grads], gn


def tf_adam_step(p, g, m, v, t, lr=0.01, b1=0.9, b2=0.999, eps=1e-4):
    """tf.train.AdamOptimizer (train.py:64). t = 1-based step count.
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t * m / (sqrt(v) + eps)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    p = p - lr_t * m / (np.sqrt(v) + eps)
    return p, m, v


# ----------------------------------------------------------------------------- synthetic data / weights
def synthetic_pianoroll(B, T, D=84, M=5, density=0.05, seed=23):
    """x = (default_rng(23).random((B,T,D,M)) < 0.05).astype(float32)."""
    rng = np.random.default_rng(seed)
    return (rng.random((B, T, D, M)) < density).astype(np.float32)


def _trunc_normal(rng, shape, std):
    """tf.truncated_normal_initializer(stddev=std): resample |z| > 2 std (common/nade.py:49-50)."""
    z = rng.standard_normal(shape)
    bad = np.abs(z) > 2
    while bad.any():
        z[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(z) > 2
    return (z * std).astype(np.float32)


def _glorot(rng, fan_in, fan_out, shape=None):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape or (fan_in, fan_out)).astype(np.float32)


def init_lstm(rng, in_dim, units):
    layers = []
    i = in_dim
    for R in units:
        layers.append((_glorot(rng, i + R, 4 * R), np.zeros((4 * R,), np.float32)))
        i = R
    return layers


def init_composer_params(D=84, M=5, H=256, R=(512, 256), seed=23):
    """Reference initialisers (SURVEY 8(d) 'Weights'), seeded."""
    rng = np.random.default_rng(seed)
    lstm = init_lstm(rng, D * M, R)
    units = M * (D + H)
    dense_p = (_glorot(rng, R[-1], units), np.zeros((units,), np.float32))
    std = 1.0 / np.sqrt(D)
    nade = [(_trunc_normal(rng, (D, H), std), _trunc_normal(rng, (D, H), std)) for _ in range(M)]
    return dict(lstm=lstm, dense=dense_p, nade=nade)


def init_rnn_nade_params(I=84, D=84, H=256, R=(512, 256), seed=23):
    rng = np.random.default_rng(seed)
    lstm = init_lstm(rng, I, R)
    dense_p = (_glorot(rng, R[-1], D + H), np.zeros((D + H,), np.float32))
    std = 1.0 / np.sqrt(D)
    nade = (_trunc_normal(rng, (D, H), std), _trunc_normal(rng, (D, H), std))
    return dict(lstm=lstm, dense=dense_p, nade=nade)


def init_rbm(rng, D, H):
    """common/rbm.py:36-58: Xavier W, zero biases."""
    return (_glorot(rng, D, H), np.zeros((1, H), np.float32), np.zeros((1, D), np.float32))


def init_rnn_rbm_params(I=84, D=84, H=256, R=(512, 256), seed=23):
    rng = np.random.default_rng(seed)
    lstm = init_lstm(rng, I, R)
    return dict(lstm=lstm, rbm=init_rbm(rng, D, H),
                Wuh=_glorot(rng, R[-1], H), Wuv=_glorot(rng, R[-1], D))


def cast_params(p, dtype):
    """Deep-cast a parameter tree to dtype."""
    if isinstance(p, dict):
        return {k: cast_params(v, dtype) for k, v in p.items()}
    if isinstance(p, (list, tuple)):
        return type(p)(cast_params(v, dtype) for v in p)
    return p.astype(dtype)


# ----------------------------------------------------------------------------- feedback modes (numpy)
def dnn_forward(x, layers):
    """common/dnn.py:97-116."""
    for K, b in layers:
        x = sigmoid(x @ K + b)
    return x


def init_dnn(rng, in_dim, units):
    layers = []
    i = in_dim
    for u in units:
        layers.append((_glorot(rng, i, u), np.zeros((u,), np.float32)))
        i = u
    return layers


def feedback_generate(xe, gen_params, fb_params, kind, num_steps, u):
    """multinn_feedback.py:120-218 on given per-track intro encodings xe[m][B,Ti+1,E] (zero-padded).
    u[S,M,B,E] uniforms. Returns sampled encodings [B,S,E,M] (before encoder.decode)."""
    M = len(xe)
    B, T1, E = xe[0].shape
    stack = np.stack(xe, axis=3).reshape(B, T1, E * M)
    if kind == 'dense':
        fb = dnn_forward(stack.reshape(B * T1, -1), fb_params).reshape(B, T1, -1)
        fb_state = None
    else:
        fb, fb_state = rnn_scan(stack, fb_params)
    states, fcs = [], []
    for m in range(M):
        p = gen_params[m]
        outs, st = rnn_scan(np.concatenate([xe[m], fb], axis=2), p['lstm'])
        states.append(st)
        fcs.append(dense(outs[:, -1], *p['dense']))
    H = gen_params[0]['nade'][0].shape[1]
    out = np.zeros((B, num_steps, E, M), xe[0].dtype)
    for s in range(num_steps):
        cur = []
        for m in range(M):
            p = gen_params[m]
            v, _ = nade_sample(fcs[m][:, :H], fcs[m][:, H:H + E], *p['nade'], u=u[s, m])
            cur.append(v)
        samples = np.stack(cur, axis=-1)                       # [B,E,M]
        out[:, s] = samples
        sstack = samples.reshape(B, E * M)
        if kind == 'dense':
            x_fb = dnn_forward(sstack, fb_params)
        else:
            x_fb, fb_state = multi_rnn_step(sstack, fb_state, fb_params)
        for m in range(M):
            p = gen_params[m]
            o, states[m] = multi_rnn_step(np.concatenate([cur[m], x_fb], axis=1), states[m], p['lstm'])
            fcs[m] = dense(o, *p['dense'])
    return out
