"""Composer MultINN (mirrors reference models/multinn/multinn_composer.py:17-199): per-track encoders, ONE
generator with a shared temporal unit and per-track NADEs (RnnMultiNADE) over the stacked encodings."""

from ..generators.rnn_multinade import RnnMultiNADE
from .core import MultINNCore


class MultINNComposer(MultINNCore):
    def __init__(self, config, params, name='MultINN-composer', **kw):
        super().__init__(config, params, name=name, **kw)
        self._mode = 'composer'

    _supports_lengths = True

    def _init_encoders(self, encoder_class):
        nh = self._params['encoder']['num_hidden']
        encs = [encoder_class(num_dims=self.num_dims, num_hidden=nh, track_name=t, arena=self._enc_arena,
                              name=f'encoder/{t}') for t in self.tracks]
        self._num_dims_generator = encs[0].num_outputs
        return encs

    def _init_generators(self, generator_class):
        if self.generator_type == 'RBM':
            raise NotImplementedError("MultiRNNRBM is not implemented yet :(")     # multinn_composer.py:44-45
        g = self._params['generator']
        self._generator = RnnMultiNADE(num_dims=self._num_dims_generator, num_hidden=g['num_hidden'],
                                       num_hidden_rnn=g['num_hidden_rnn'], tracks=self.tracks,
                                       keep_prob=self.keep_prob, arena=self._arena, name='generator')
        return [self._generator]

    def _forward_backward(self, x, keep, u_drop, seed, lengths=None, loss_scale=1.0, u_enc=None, **extra):
        B, T, D, M = x.shape
        # per-track encodings (PassEncoder: the piano-rolls themselves; DBNEncoder: sampled codes,
        # core/multi_encoder_nn.py:98-115) stacked with feature e*M + m (multinn_composer.py:73-80)
        _, stack, bits = self._encode_tracks(x, u_enc, seed, need_tracks=False)
        # inputs = slots 0..T-1 ([0, x_0..x_{T-2}]), targets = slots 1..T (multinn_composer.py:82-86)
        loss, nll, _ = self._generator.forward_backward(stack[:T], bits, keep=keep, u_drop=u_drop, seed=seed,
                                                        lengths=lengths, loss_scale=loss_scale)
        self._last_nll = (nll, T, B)
        return loss

    def evaluate(self, x, lengths=None, cond_probs=False, u_enc=None, seed=0):
        """is_train=False forward: per-row NLL[N,M] (rows n = b*T + t), `batch/loss` = mean over tracks of the
        per-track means (metrics/statistical.py:34, rnn_multinade.py:200-203). With variable `lengths` the rows
        t >= lengths[b] are removed (utils/sequences.py:29-37): `nll` then holds the valid rows only, b-major."""
        x = self._check_x(x, lengths)
        B, T, D, M = x.shape
        _, stack, bits = self._encode_tracks(x, u_enc, seed, need_tracks=False)
        nll, cp = self._generator.log_prob(stack[:T], bits, cond_probs=cond_probs, lengths=lengths)
        D = self._num_dims_generator
        keep_rows = self.valid_rows(lengths, T, B, x.device)
        out = {'nll': self.rows_to_reference_order(nll, T, B)}
        if cp is not None:
            out['cond_probs'] = cp.view(M, T, B, D).permute(2, 1, 3, 0).reshape(B * T, D, M)
        if keep_rows is not None:
            out = {k: v[keep_rows] for k, v in out.items()}
        out['batch/loss'] = out['log_likelihood'] = out['nll'].mean(0).mean()
        self._metrics.update(out)
        return out

    def generate(self, x, num_steps, u=None, seed=0, u_enc=None, u_dec=None):
        """multinn_composer.py:114-151. x[B,Ti,D,M] intro -> samples[B,num_steps,D,M]; u[num_steps,M,B,E]. DBN encoders:
        the intro is encoded per track (u_enc), the generated codes are decoded per track (u_dec, :140-150)."""
        x = self._check_x(x, None)
        B, T, D, M = x.shape
        _, stack, _ = self._encode_tracks(x, u_enc, seed, need_tracks=False)
        samples = self._generator.generate(stack, num_steps, u=u, seed=seed)     # whole padded intro
        return self._decode_tracks(samples.view(B, num_steps, self._num_dims_generator, M), u_dec, seed)
