"""RNN-RBM generator (mirrors reference models/generators/rnn_rbm.py:17-322): LSTM temporal unit whose outputs set the
per-step RBM biases bh_t = u_t Wuh (+ bh), bv_t = u_t Wuv (+ bv) (internal_bias=True by default, :22, :240-259); samples
come from a k-step Gibbs chain STARTED FROM THE INPUT FRAME (:112, :295).

Training loss as written in the reference (quirk Q3): free-energy cost with the RBM's INTERNAL biases only, on the
stop-gradient chain sample, so the gradient reaches W, bh, bv only; LSTM / Wuh / Wuv receive no gradient (their grads stay
zero here, which TF-Adam turns into a zero update like the reference's None gradients).
`conditional_free_energy=True` is the canonical RNN-RBM cost with bh_t, bv_t (opt-in, off for parity).
"""
import torch

from .. import ops
from ..common.rbm import RBM
from ..params import glorot_uniform
from .rnn_estimator import RnnEstimator, RnnEstimatorStateTuple


class RnnRBM(RnnEstimator):
    def __init__(self, num_dims, num_hidden, num_hidden_rnn, keep_prob=1.0, internal_bias=True, k=10, name='rnn-rbm',
                 track_name='all', arena=None, num_inputs=None, conditional_free_energy=False):
        self._k = k
        if conditional_free_energy:
            raise NotImplementedError('conditional_free_energy=True (canonical RNN-RBM cost) is not implemented yet')
        super().__init__(arena, num_inputs if num_inputs is not None else num_dims, num_dims, num_hidden,
                         num_hidden_rnn, keep_prob, internal_bias, name, track_name, binary_inputs=num_inputs is None)

    def _init_estimator(self):
        D, H, r_top = self._num_dims, self._num_hidden[-1], self._num_hidden_rnn[-1]
        self._rbm = RBM(D, H, k=self._k, name=f'{self.name}/rbm', arena=self._arena)              # rnn_rbm.py:52
        self._Wuh = self._arena.add(f'{self.name}/Wuh', (r_top, H), glorot_uniform(r_top, H))     # :54-58
        self._Wuv = self._arena.add(f'{self.name}/Wuv', (r_top, D), glorot_uniform(r_top, D))     # :60-64

    rbm = property(lambda s: s._rbm)
    num_outputs = property(lambda s: s._num_dims)
    k = property(lambda s: s._k)

    @property
    def trainable_params(self):
        return self._rbm.trainable_params + self._rnn.trainable_params + [self._Wuh, self._Wuv]

    def zero_state(self, batch_size, device='cuda'):
        D, H = self._num_dims, self._num_hidden[-1]
        return RnnEstimatorStateTuple(torch.zeros(batch_size, H, device=device), torch.zeros(batch_size, D, device=device),
                                      self._rnn.zero_state(batch_size, device))

    def _build_biases(self, outputs):
        """rnn_rbm.py:240-259."""
        N = outputs.shape[0]
        bh = torch.empty(N, self._num_hidden[-1], device=outputs.device)
        bv = torch.empty(N, self._num_dims, device=outputs.device)
        ops.gemm(outputs, self._Wuh.data, bh, bias=self._rbm.bh.data.view(-1) if self._internal_bias else None)
        ops.gemm(outputs, self._Wuv.data, bv, bias=self._rbm.bv.data.view(-1) if self._internal_bias else None)
        return bh, bv

    def _get_state(self, inputs, lengths=None, initial_state=None, last_outputs=False, keep=1.0, u_drop=None, seed=0):
        """rnn_rbm.py:184-238 (tf.nn.dynamic_rnn, full lengths). inputs[T,B,I] time-major."""
        if lengths is not None:
            raise NotImplementedError('variable `lengths` is a next-row item; pass None')
        if inputs.dim() == 2:
            inputs = inputs.unsqueeze(0)
        T, B, _ = inputs.shape
        rnn_init = None if initial_state is None else initial_state.rnn_state
        outs, rnn_state = self._rnn.forward_sequence(inputs.contiguous(), keep=keep, u=u_drop, seed=seed,
                                                     initial_state=rnn_init)
        if last_outputs:
            flat = outs[T - 1]
            rnn_state = [type(s)(s[0].clone(), s[1].clone()) for s in rnn_state]
        else:
            flat = outs.reshape(T * B, -1)
        bh_t, bv_t = self._build_biases(flat)
        return RnnEstimatorStateTuple(bh_t, bv_t, rnn_state)

    def single_step(self, inputs, initial_state):
        """rnn_rbm.py:261-281."""
        out, rnn_state = self._rnn.step(inputs, initial_state.rnn_state)
        bh_t, bv_t = self._build_biases(out)
        return RnnEstimatorStateTuple(bh_t, bv_t, rnn_state)

    def sample_single(self, inputs, state, u=None, seed=0, offset=0, out=None, temperature=None):
        """rnn_rbm.py:283-297: (sample, cond_prob) of a k-step chain started at `inputs`. u = (uh[k,B,H], uv[k,B,D])."""
        cond_prob, sample = self._rbm.sample(inputs.contiguous(), state.b_enc, state.b_dec, k=self._k, u=u,
                                             seed=seed * 7919 + offset)
        if out is not None:
            out.copy_(sample)
            sample = out
        return sample, cond_prob

    def pretrain(self, inputs_flat, lr, u=None, seed=None):
        """rnn_rbm.py:299-322: pre-trains the RBM module with one CD-k update on the flattened input frames (the LSTM,
        Wuh and Wuv are untouched). u = dict(uh, uv, uh0, uhk) uniforms for parity runs. Returns (p_v, v_k) of the chain."""
        return self._rbm.train(inputs_flat.contiguous(), lr, u=u, seed=seed)

    # ------------------------------------------------------------------ train / eval graph (rnn_rbm.py:94-119)
    def forward(self, inputs, targets, keep=1.0, u_drop=None, u_gibbs=None, seed=0):
        """inputs[T,B,I] (I == num_dims: the chain starts from the input frame), targets[T,B,D] time-major.
        Returns dict(loss[1] = mean free-energy cost, free_energy[1], sample[N,D], cond_probs[N,D]); rows n' = t*B + b."""
        T, B, _ = inputs.shape
        state = self._get_state(inputs, keep=keep, u_drop=u_drop, seed=seed)
        v0 = inputs.reshape(T * B, -1)
        sample, cond_prob = self.sample_single(v0, state, u=u_gibbs, seed=seed)
        tgt = targets.reshape(T * B, -1)
        cost, fe = self._rbm.free_energy_cost(tgt, sample)
        self._saved_fb = (tgt, sample)
        return dict(loss=cost, free_energy=fe, sample=sample, cond_probs=cond_prob)

    def forward_backward(self, inputs, targets, keep=None, u_drop=None, u_gibbs=None, seed=0, loss_scale=1.0):
        keep = self._keep_prob if keep is None else keep
        out = self.forward(inputs, targets, keep=keep, u_drop=u_drop, u_gibbs=u_gibbs, seed=seed)
        tgt, sample = self._saved_fb
        self._rbm.free_energy_cost_backward(tgt, sample, scale=loss_scale)
        return out['loss'] * loss_scale if loss_scale != 1.0 else out['loss'], out
