"""Jamming MultINN (mirrors reference models/multinn/multinn_jamming.py:16-245): M independent per-track
generators; the generators are optimised JOINTLY: one loss = mean of track losses, one global-norm clip and
one Adam over the union of all generators' variables (:235-243, quirk Q7)."""
import torch

from .core import MultINNCore


class MultINNJamming(MultINNCore):
    def __init__(self, config, params, name='MultINN-jamming', **kw):
        super().__init__(config, params, name=name, **kw)
        self._mode = 'jamming'

    _supports_lengths = True

    def _init_encoders(self, encoder_class):
        nh = self._params['encoder']['num_hidden']
        encs = [encoder_class(num_dims=self.num_dims, num_hidden=nh, track_name=t, arena=self._enc_arena,
                              name=f'encoder/{t}') for t in self.tracks]
        self._num_dims_generator = encs[0].num_outputs
        return encs

    def _init_generators(self, generator_class):
        g = self._params['generator']
        return [generator_class(num_dims=self._num_dims_generator, num_hidden=g['num_hidden'],
                                num_hidden_rnn=g['num_hidden_rnn'], keep_prob=self.keep_prob, track_name=t,
                                arena=self._arena, name=f'generator/{t}') for t in self.tracks]

    def _forward_backward(self, x, keep, u_drop, seed, lengths=None, loss_scale=1.0, u_enc=None, **extra):
        B, T, D, M = x.shape
        xe, _, bits = self._encode_tracks(x, u_enc, seed, need_stack=False)          # multinn_jamming.py:60-68 over the track encodings
        def one(m, gen):
            return gen.forward_backward(xe[m][:T], bits[m:m + 1], keep=keep, u_drop=None if u_drop is None else u_drop[m],
                                        seed=seed + 104729 * m, loss_scale=loss_scale / M, lengths=lengths)
        total = torch.zeros(1, device=x.device)
        for loss, _, _ in self._per_track(one):
            total += loss
        return total

    def evaluate(self, x, lengths=None, u_enc=None, seed=0):
        x = self._check_x(x, lengths)
        B, T, D, M = x.shape
        xe, _, bits = self._encode_tracks(x, u_enc, seed, need_stack=False)
        nll = torch.empty(M, T * B, device=x.device)
        for m, gen in enumerate(self._generators):
            n, _ = gen.log_prob(xe[m][:T], bits[m:m + 1], lengths=lengths)
            nll[m] = n[0]
        out = {'nll': self.rows_to_reference_order(nll, T, B)}
        keep_rows = self.valid_rows(lengths, T, B, x.device)
        if keep_rows is not None:
            out['nll'] = out['nll'][keep_rows]
        out['batch/loss'] = out['log_likelihood'] = out['nll'].mean(0).mean()     # multinn_core.py:402-405
        self._metrics.update(out)
        return out

    def generate(self, x, num_steps, u=None, seed=0, u_enc=None, u_dec=None):
        """multinn_jamming.py:101-133: every track generated independently over its encodings, then decoded;
        u[num_steps,M,B,E]."""
        x = self._check_x(x, None)
        B, T, D, M = x.shape
        xe, _, _ = self._encode_tracks(x, u_enc, seed, need_stack=False)
        out = torch.empty(B, num_steps, self._num_dims_generator, M, device=x.device)
        for m, gen in enumerate(self._generators):
            s = gen.generate(xe[m].contiguous(), num_steps, u=None if u is None else u[:, m:m + 1], seed=seed + 104729 * m)
            out[..., m] = s
        return self._decode_tracks(out, u_dec, seed)
