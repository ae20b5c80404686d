"""Joint MultINN (mirrors reference models/multinn/multinn_joint.py:17-215): ONE encoder over all tracks stacked on the
feature axis (D*M dims, feature d*M + m) and ONE generator over its encodings."""
import torch

from ..generators.rnn_rbm import RnnRBM
from .core import MultINNCore


class MultINNJoint(MultINNCore):
    def __init__(self, config, params, name='MultINN-joint', **kw):
        super().__init__(config, params, name=name, **kw)
        self._mode = 'joint'

    def _init_encoders(self, encoder_class):
        nh = self._params['encoder']['num_hidden']
        self._encoder = encoder_class(num_dims=self.num_dims * self.num_tracks, num_hidden=nh, track_name='all',
                                      arena=self._enc_arena, name='encoder/all')
        self._num_dims_generator = self._encoder.num_outputs
        return [self._encoder]

    def _init_generators(self, generator_class):
        g = self._params['generator']
        cls = RnnRBM if generator_class == 'RBM' else generator_class
        self._generator = cls(num_dims=self._num_dims_generator, num_hidden=g['num_hidden'],
                              num_hidden_rnn=g['num_hidden_rnn'], keep_prob=self.keep_prob, arena=self._arena,
                              name='generator')
        return [self._generator]

    def _encode(self, x, u_enc=None, seed=0):
        """_build_inputs (:76-89: reshape to [B,T,D*M], zero-pad the front) + encoder.encode() (sampled codes, quirk Q12).
        Returns time-major codes[(T+1),B,E]."""
        B, T, D, M = x.shape
        st = self._stage_inputs(x, stacked=True)
        flat = st['xin'].view((T + 1) * B, D * M)
        _, h = self._encoder.encode(flat, u=u_enc, seed=seed)
        return h.view(T + 1, B, -1)

    def _pretrain_rows(self, x, seed):
        """The generator's input frames: the codes of the zero-padded sequence without its last step, flattened."""
        if not isinstance(self._generator, RnnRBM):
            return [(self._generator, None)]
        B, T, D, M = x.shape
        codes = self._encode(x, None, seed)
        return [(self._generator, codes[:T].reshape(T * B, -1))]

    def _forward_backward(self, x, keep, u_drop, seed, u_enc=None, u_gibbs=None):
        B, T, D, M = x.shape
        codes = self._encode(x, u_enc, seed)
        if isinstance(self._generator, RnnRBM):
            # train_generators (:263-288) optimises the GENERATOR's own batch/loss (generator.py:196-201): no 1/M;
            # :182-184 divides only the model-level (encoder `global`) metrics by M.
            loss, out = self._generator.forward_backward(codes[:T], codes[1:], keep=keep, u_drop=u_drop, u_gibbs=u_gibbs,
                                                         seed=seed)
            return loss
        bits = torch.empty(1, T * B, 4, dtype=torch.int32, device=x.device)
        from .. import ops
        ops.pack_rows(codes[1:].reshape(T * B, -1), bits[0])
        loss, nll, _ = self._generator.forward_backward(codes[:T], bits, keep=keep, u_drop=u_drop, seed=seed)
        return loss

    def evaluate(self, x, lengths=None, u_enc=None, u_gibbs=None, seed=0):
        """is_train=False forward: the generator's own metrics (what train.py:75-76 monitors). RBM generator:
        `batch/loss` = mean free-energy cost, `nll` = per-row sum of tf.losses.log_loss(targets, cond_probs) (eps 1e-7,
        common/rbm.py:121-129) in the reference's row order, `sample` = chain samples."""
        x = self._check_x(x, lengths)
        B, T, D, M = x.shape
        codes = self._encode(x, u_enc, seed)
        if isinstance(self._generator, RnnRBM):
            out = self._generator.forward(codes[:T], codes[1:], u_gibbs=u_gibbs, seed=seed)
            t, p = codes[1:].reshape(T * B, -1), out['cond_probs']
            nll = -(t * torch.log(p + 1e-7) + (1 - t) * torch.log(1 - p + 1e-7)).sum(1)
            nll_ref = self.rows_to_reference_order(nll.view(1, -1), T, B)
            res = {'nll': nll_ref, 'log_likelihood': nll_ref.mean(), 'batch/loss': out['loss'],
                   'free_energy': out['free_energy'], 'codes': codes, 'sample': out['sample'],
                   'cond_probs': out['cond_probs']}
        else:
            from .. import ops
            bits = torch.empty(1, T * B, 4, dtype=torch.int32, device=x.device)
            ops.pack_rows(codes[1:].reshape(T * B, -1), bits[0])
            nll, _ = self._generator.log_prob(codes[:T], bits)
            nll_ref = self.rows_to_reference_order(nll, T, B)
            res = {'nll': nll_ref, 'log_likelihood': nll_ref.mean(), 'batch/loss': nll_ref.mean(), 'codes': codes}
        self._metrics.update(res)
        return res

    def generate(self, x, num_steps, u=None, seed=0, u_enc=None, u_dec=None):
        """multinn_joint.py:188-215: intro codes -> generator.generate -> encoder.decode -> [B,S,D,M].
        u: per-step uniforms for the generator (RBM: list of (uh[k,B,H], uv[k,B,D]); NADE: [S,1,B,E])."""
        x = self._check_x(x, None)
        B, T, D, M = x.shape
        codes = self._encode(x, u_enc, seed)
        samples_h = self._generator.generate(codes, num_steps, u=u, seed=seed)          # [B,S,E]
        from .. import ops
        with ops.row_map_scaled(num_steps):            # decode rows are b-major (b*S + s)
            _, v = self._encoder.decode(samples_h.reshape(B * num_steps, -1), u=u_dec, seed=seed + 1)
        return v.view(B, num_steps, D, M)
