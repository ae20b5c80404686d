"""Feedback MultINN (mirrors reference models/multinn/multinn_feedback.py:17-218): per-track encoders and generators;
every generator's input is [its own track's encoding ; feedback(stacked encodings of ALL tracks)] where the feedback
module is a sigmoid Dense net (here) or an LSTM stack (feedback_rnn.py). Generators and feedback module are optimised
jointly with one global-norm clip (multinn_jamming.py:235-243 via inheritance, quirk Q7)."""
import torch

from .. import ops
from ..common.dnn import DNN
from ..generators.rnn_nade import RnnNade
from .core import MultINNCore


class MultINNFeedback(MultINNCore):
    def __init__(self, config, params, name='MultINN-feedback', **kw):
        super().__init__(config, params, name=name, **kw)
        self._mode = 'feedback'
        self._feedback_module = True

    # ------------------------------------------------------------------ construction
    def _init_encoders(self, encoder_class):
        nh = self._params['encoder']['num_hidden']
        encs = [encoder_class(num_dims=self.num_dims, num_hidden=nh, track_name=t, arena=self._enc_arena,
                              name=f'encoder/{t}') for t in self.tracks]
        self._num_dims_generator = encs[0].num_outputs
        return encs

    def _feedback_units(self):
        fb = self._params['generator']['feedback']
        if fb is None:
            raise ValueError('feedback modes need `generator.feedback`, e.g. [128, 128]')
        return [fb] if isinstance(fb, int) else list(fb)

    def _init_feedback(self):
        """multinn_feedback.py:46-52."""
        return DNN(self._arena, self._num_dims_generator * self.num_tracks, self._feedback_units(), name='feedback')

    def _init_generators(self, generator_class):
        if generator_class == 'RBM':
            raise NotImplementedError('RnnRBM generators cannot take [encoding ; feedback] inputs (the Gibbs chain '
                                      'starts from the input frame, generators/rnn_rbm.py:112)')
        g = self._params['generator']
        E, F = self._num_dims_generator, self._feedback_units()[-1]
        gens = [RnnNade(num_dims=E, num_hidden=g['num_hidden'], num_hidden_rnn=g['num_hidden_rnn'],
                        keep_prob=self.keep_prob, track_name=t, arena=self._arena, name=f'generator/{t}',
                        num_inputs=E + F) for t in self.tracks]
        self._feedback_layer = self._init_feedback()           # after the generators, like the variable order in TF
        return gens

    # ------------------------------------------------------------------ encodings
    def _encode(self, x, u_enc=None, seed=0):
        """Per-track encodings, their stack (feature e*M + m, multinn_feedback.py:67-73) and target bit masks."""
        return self._encode_tracks(x, u_enc, seed)

    def _apply_feedback(self, stack, keep=1.0, u_fb=None, seed=0, save=True):
        """multinn_feedback.py:99-118: Dense feedback over every (padded) step. stack[(T+1),B,E*M] -> [(T+1),B,F]."""
        T1, B, _ = stack.shape
        return self._feedback_layer(stack.reshape(T1 * B, -1), save=save).view(T1, B, -1)

    def _feedback_backward(self, dfb):
        T1, B, F = dfb.shape
        self._feedback_layer.backward(dfb.reshape(T1 * B, F))

    # ------------------------------------------------------------------ train / eval
    _supports_lengths = True      # the generators take `lengths` (multinn_feedback.py:93-94); the feedback module runs
                                  # over every padded step in training (:76-80 passes no lengths)

    def _forward_backward(self, x, keep, u_drop, seed, u_enc=None, u_fb=None, lengths=None, loss_scale=1.0, **extra):
        B, T, D, M = x.shape
        xe, stack, bits = self._encode(x, u_enc, seed)
        fb = self._apply_feedback(stack, keep=keep, u_fb=u_fb, seed=seed + 17)
        E, F = self._num_dims_generator, fb.shape[2]
        dfb = torch.zeros(T + 1, B, F, device=x.device)
        def one(m, gen):
            inp = torch.cat([xe[m][:T], fb[:T]], dim=2)                       # multinn_feedback.py:86-88
            return gen.forward_backward(inp, bits[m:m + 1], keep=keep, u_drop=None if u_drop is None else u_drop[m],
                                        seed=seed + 104729 * m, loss_scale=loss_scale / M, need_dx=True, lengths=lengths)
        total = torch.zeros(1, device=x.device)
        for loss, _, dx in self._per_track(one):
            total += loss
            dfb[:T] += dx[:, :, E:]
        self._feedback_backward(dfb)
        return total

    def evaluate(self, x, lengths=None, u_enc=None, seed=0):
        x = self._check_x(x, lengths)
        B, T, D, M = x.shape
        xe, stack, bits = self._encode(x, u_enc, seed)
        fb = self._apply_feedback(stack, save=False)
        nll = torch.empty(M, T * B, device=x.device)
        for m, gen in enumerate(self._generators):
            n, _ = gen.log_prob(torch.cat([xe[m][:T], fb[:T]], dim=2), bits[m:m + 1], lengths=lengths)
            nll[m] = n[0]
        out = {'nll': self.rows_to_reference_order(nll, T, B)}
        keep_rows = self.valid_rows(lengths, T, B, x.device)
        if keep_rows is not None:
            out['nll'] = out['nll'][keep_rows]
        out['batch/loss'] = out['log_likelihood'] = out['nll'].mean(0).mean()
        self._metrics.update(out)
        return out

    # ------------------------------------------------------------------ generation (multinn_feedback.py:120-218)
    def _feedback_step(self, samples_stack, state):
        return self._feedback_layer(samples_stack, save=False), state

    def _feedback_intro_state(self):
        return None

    def generate(self, x, num_steps, u=None, seed=0, u_enc=None, u_dec=None):
        """x[B,Ti,D,M] -> [B,num_steps,D,M]; u[num_steps,M,B,E] uniforms for the NADE samplers."""
        x = self._check_x(x, None)
        B, T, D, M = x.shape
        E = self._num_dims_generator
        xe, stack, _ = self._encode(x, u_enc, seed)
        fb = self._apply_feedback(stack, save=False)
        fb_state = self._feedback_intro_state()
        states = [gen.steps(torch.cat([xe[m], fb], dim=2)) for m, gen in enumerate(self._generators)]
        samples_h = torch.empty(B, num_steps, E, M, device=x.device)
        prev = None
        for s in range(num_steps):
            cur = torch.empty(B, E, M, device=x.device)
            for m, gen in enumerate(self._generators):
                tmp = torch.empty(B, E, device=x.device)
                gen.sample_single(prev, states[m], u=None if u is None else u[s, m:m + 1], seed=seed + 104729 * m,
                                  offset=s, out=tmp)
                cur[:, :, m] = tmp
            samples_h[:, s] = cur
            x_fb, fb_state = self._feedback_step(cur.reshape(B, E * M), fb_state)
            states = [gen.single_step(torch.cat([cur[:, :, m], x_fb], dim=1), states[m])
                      for m, gen in enumerate(self._generators)]
            prev = cur
        music = self._decode_tracks(samples_h, u_dec, seed)
        return music
