"""RNN-NADE generator (mirrors reference models/generators/rnn_nade.py:21-326) generalised over M tracks so that
RnnMultiNADE (rnn_multinade.py) is the same code path with M > 1.

Pipeline per sequence batch (time-major, rows n' = t*B + b):
  LSTM stack -> Dense `fc[N, M*(H+D)]` (rnn_nade.py:54-57, :212) -> column split [M*H | M*D]
  (rnn_nade.py:234-251, rnn_multinade.py:231-256; the NADE kernels read the columns in place) ->
  teacher-forced NADE NLL per track (common/nade.py:155-229).
Training = fwd + analytic backward of the same graph tf.gradients would differentiate (generator.py:176-205).
"""
import torch

from .. import ops
from ..common.nade import NADE, NADEBank
from ..params import glorot_uniform, zeros
from .rnn_estimator import RnnEstimator, RnnEstimatorStateTuple


class RnnNade(RnnEstimator):
    def __init__(self, num_dims, num_hidden, num_hidden_rnn, keep_prob=1.0, internal_bias=False, name='rnn-nade',
                 track_name='all', arena=None, num_inputs=None, num_tracks=1):
        if internal_bias:
            raise NotImplementedError('internal_bias=True is not used by any MultINN mode for NADE generators')
        self._num_tracks = num_tracks
        super().__init__(arena, num_inputs if num_inputs is not None else num_dims * num_tracks, num_dims,
                         num_hidden, num_hidden_rnn, keep_prob, internal_bias, name, track_name,
                         binary_inputs=num_inputs is None)   # own-track / stacked encodings are exactly {0,1}
        self._ws = {}
        self._saved = None

    # -------------------------------------------------------------- construction
    def _init_estimator(self):
        M, D, H = self._num_tracks, self._num_dims, self._num_hidden[-1]
        r_top = self._num_hidden_rnn[-1]
        units = M * (D + H)
        self._fc_kernel = self._arena.add(f'{self.name}/dense/kernel', (r_top, units), glorot_uniform(r_top, units))
        self._fc_bias = self._arena.add(f'{self.name}/dense/bias', (units,), zeros())
        self._bank = NADEBank(self._arena, M, D, H, name=f'{self.name}/nade')
        self._nades = [NADE(D, H, bank=self._bank, track=m, name=f'{self.name}/nade{m}') for m in range(M)]

    num_tracks = property(lambda s: s._num_tracks)
    num_outputs = property(lambda s: s._num_tracks * s._num_dims)
    enc_col0 = 0
    dec_col0 = property(lambda s: s._num_tracks * s._num_hidden[-1])

    @property
    def trainable_params(self):
        return self._rnn.trainable_params + [self._fc_kernel, self._fc_bias] + self._bank.trainable_params

    def zero_state(self, batch_size, device='cuda'):
        """generator.py zero_state; rnn_multinade.py:223-229 (quirk Q6 implemented as intended)."""
        M, D, H = self._num_tracks, self._num_dims, self._num_hidden[-1]
        fc = torch.zeros(batch_size, M * (H + D), device=device)
        return self._state_from_fc(fc, self._rnn.zero_state(batch_size, device))

    # -------------------------------------------------------------- state helpers
    def _state_from_fc(self, fc, rnn_state):
        """_build_biases: track m's b_enc = cols [m*H,(m+1)*H), b_dec = cols [M*H + m*D, ...)."""
        M, D, H = self._num_tracks, self._num_dims, self._num_hidden[-1]
        be = [fc[:, m * H:(m + 1) * H] for m in range(M)]
        bd = [fc[:, M * H + m * D:M * H + (m + 1) * D] for m in range(M)]
        st = RnnEstimatorStateTuple(be if M > 1 else be[0], bd if M > 1 else bd[0], rnn_state)
        self._last_fc = fc
        return st

    def _fc_of(self, state):
        be0 = state.b_enc[0] if isinstance(state.b_enc, (list, tuple)) else state.b_enc
        # b_enc of track 0 starts at column 0 of the Dense buffer it is a view of
        M, D, H = self._num_tracks, self._num_dims, self._num_hidden[-1]
        return be0.as_strided((be0.shape[0], M * (H + D)), (be0.stride(0), 1))

    def _workspace(self, N, device, training):
        key = (N, training)
        ws = self._ws.get(key)
        if ws is None:
            M, D, H = self._num_tracks, self._num_dims, self._num_hidden[-1]
            U = M * (H + D)
            ws = dict(fc=torch.empty(N, U, device=device), nll=torch.empty(M, N, device=device),
                      loss=torch.zeros(1, device=device))
            if training:
                ws['dfc'] = torch.empty(N, U, device=device)
                ws['dout'] = torch.empty(N, self._num_hidden_rnn[-1], device=device)
            self._ws = {key: ws}
        return ws

    def _get_state(self, inputs, lengths=None, initial_state=None, last_outputs=False, keep=1.0, u_drop=None,
                   seed=0, training=False):
        """rnn_nade.py:173-232. inputs[T,B,I] time-major (or [B,I] = one step). dynamic_decode(impute_finished=False)
        computes the rows past a sequence's length too; `log_prob` / `forward_backward` take `lengths` and
        give those rows weight 0 (flatten_maybe_padded_sequences drops them, utils/sequences.py:6-37)."""
        if inputs.dim() == 2:
            inputs = inputs.unsqueeze(0)
        T, B, _ = inputs.shape
        rnn_init = None if initial_state is None else initial_state.rnn_state
        outs, rnn_state = self._rnn.forward_sequence(inputs.contiguous(), keep=keep, u=u_drop, seed=seed,
                                                     initial_state=rnn_init)
        if last_outputs:
            fc = torch.empty(B, self._fc_kernel.shape[1], device=inputs.device)
            ops.gemm(outs[T - 1], self._fc_kernel.data, fc, bias=self._fc_bias.data)
            rnn_state = [type(s)(s[0].clone(), s[1].clone()) for s in rnn_state]
        else:
            ws = self._workspace(T * B, inputs.device, training)
            fc = ws['fc']
            ops.gemm(outs.reshape(T * B, -1), self._fc_kernel.data, fc, bias=self._fc_bias.data, b_weight=True)
        self._outs = outs
        return self._state_from_fc(fc, rnn_state)

    # -------------------------------------------------------------- teacher-forced likelihood
    @staticmethod
    def row_weights(lengths, T, B, device):
        """w[t*B + b] = 1 if t < lengths[b] else 0 (tf.sequence_mask, utils/sequences.py:29), and the number of valid
        rows; None for full lengths (the reshape branch, sequences.py:22-24)."""
        if lengths is None:
            return None, T * B
        lengths = torch.as_tensor(lengths).to('cpu', torch.int64)
        if lengths.numel() != B:
            raise ValueError(f'lengths must hold one entry per sequence ({B}), got {lengths.numel()}')
        if int(lengths.min()) >= T:
            return None, T * B
        if int(lengths.min()) < 0:
            raise ValueError('negative sequence length')
        w = (torch.arange(T)[:, None] < lengths[None, :]).to(torch.float32).reshape(-1)
        return w.to(device), int(lengths.clamp(max=T).sum())

    def log_prob(self, inputs, bits, keep=1.0, u_drop=None, seed=0, cond_probs=False, lengths=None):
        """rnn_nade.py:279-302 / rnn_multinade.py:258-290. inputs[T,B,I], bits[M,T*B,4] target masks.
        Returns (nll[M,N'] positive, cond_p[M,N',D] or None); rows n' = t*B + b. With `lengths`, the NLL of the rows
        t >= lengths[b] is 0 (the reference removes those rows)."""
        T, B, _ = inputs.shape
        w, _ = self.row_weights(lengths, T, B, inputs.device)
        self._get_state(inputs, keep=keep, u_drop=u_drop, seed=seed)
        ws = self._ws[(T * B, False)]
        cp = None
        if cond_probs:
            cp = torch.empty(self._num_tracks, T * B, self._num_dims, device=inputs.device)
        ops.nade_logprob_fwd(bits, ws['fc'], self.enc_col0, self.dec_col0, self._bank.w_enc.data,
                             self._bank.w_dec.data, ws['nll'], cond_p=cp)
        if w is not None:
            ops.scale_rows(ws['nll'].view(-1, 1), w, period=T * B)
        return ws['nll'], cp

    def forward_backward(self, inputs, bits, keep=None, u_drop=None, seed=0, loss_scale=1.0, need_dx=False,
                         lengths=None):
        """One training pass: loss = loss_scale * mean_m mean_n NLL (statistical.py:34; rnn_multinade.py:200-203)
        and its gradient wrt every trainable parameter (written into the arena's grad buffer; NADE weight grads
        are ACCUMULATED, so the caller zeroes the grad buffer once per step). Returns (loss[1], nll[M,N'], dx).
        `lengths`: the means run over the valid rows only; padded rows get zero NLL and zero logit gradient."""
        keep = self._keep_prob if keep is None else keep
        T, B, _ = inputs.shape
        N, M = T * B, self._num_tracks
        w, nvalid = self.row_weights(lengths, T, B, inputs.device)
        if w is None and self._rnn.use_any_pipeline(T, B):
            return self._forward_backward_pipelined(inputs, bits, keep, u_drop, seed, loss_scale, need_dx)
        self._get_state(inputs, keep=keep, u_drop=u_drop, seed=seed, training=True)
        ws = self._ws[(N, True)]
        gscale = loss_scale / (nvalid * M)
        ops.nade_logprob_fwd(bits, ws['fc'], self.enc_col0, self.dec_col0, self._bank.w_enc.data,
                             self._bank.w_dec.data, ws['nll'], dfc=ws['dfc'], gscale=gscale)
        if w is not None:       # every NADE gradient is linear in the logit gradient the forward wrote into dfc's b_dec columns
            ops.scale_rows(ws['nll'].view(-1, 1), w, period=N)
            ops.scale_rows(ws['dfc'][:, self.dec_col0:], w)
        ops.sum_into(ws['nll'], ws['loss'], scale=gscale)
        ops.nade_logprob_bwd(bits, ws['fc'], self.enc_col0, self.dec_col0, self._bank.w_enc.data,
                             self._bank.w_dec.data, ws['dfc'], self._bank.w_enc.grad, self._bank.w_dec.grad)
        outs = self._outs.reshape(N, -1)
        ops.gemm(outs, ws['dfc'], self._fc_kernel.grad, transA=True)
        ops.colsum(ws['dfc'], self._fc_bias.grad)
        ops.gemm(ws['dfc'], self._fc_kernel.data, ws['dout'], transB=True, b_weight=True)
        dx = self._rnn.backward_sequence(ws['dout'].view(T, B, -1), need_dx=need_dx)
        return ws['loss'], ws['nll'], dx

    def _forward_backward_pipelined(self, inputs, bits, keep, u_drop, seed, loss_scale, need_dx):
        """Same results as the phase-by-phase path, scheduled as a pipeline over time chunks (small per-GPU batches, where
        the recurrences are latency chains that leave SMs idle): as soon as the top LSTM layer has finished a chunk, its
        Dense forward, NADE forward + backward, Dense weight/bias gradients (accumulated) and the gradient wrt the LSTM
        outputs run on the bulk stream beside the recurrences of the later chunks; BPTT then overlaps the LSTM weight
        gradients of the chunks it has left behind (common/rnn.py)."""
        T, B, _ = inputs.shape
        N, M = T * B, self._num_tracks
        rnn = self._rnn
        ws = self._workspace(N, inputs.device, True)
        gscale = loss_scale / (N * M)
        fc, dfc, nll, dout = ws['fc'], ws['dfc'], ws['nll'], ws['dout']
        bulk = rnn.bulk_stream(inputs.device)
        main = torch.cuda.current_stream()
        ready = rnn._event('step start')
        ready.record(main)                   # zeroed gradient buffers, staged inputs
        bulk.wait_event(ready)
        last = [None]
        C = rnn.WAVEFRONT_CHUNK
        dout_ready = {}                      # chunk -> events after which dout of that chunk is final (bulk stream)

        def hook(t0, t1, done, outs_top, budget, is_last):
            r0, r1 = t0 * B, t1 * B
            oc = outs_top[t0:t1].reshape(r1 - r0, -1)
            with torch.cuda.stream(bulk):
                bulk.wait_event(done)
                ops.set_sm_budget(budget)
                try:
                    ops.gemm(oc, self._fc_kernel.data, fc[r0:r1], bias=self._fc_bias.data, b_weight=True)
                    ops.nade_logprob_fwd(bits[:, r0:r1], fc[r0:r1], self.enc_col0, self.dec_col0, self._bank.w_enc.data,
                                         self._bank.w_dec.data, nll[:, r0:r1], dfc=dfc[r0:r1], gscale=gscale)
                    ops.nade_logprob_bwd(bits[:, r0:r1], fc[r0:r1], self.enc_col0, self.dec_col0, self._bank.w_enc.data,
                                         self._bank.w_dec.data, dfc[r0:r1], self._bank.w_enc.grad, self._bank.w_dec.grad)
                    ops.gemm(dfc[r0:r1], self._fc_kernel.data, dout[r0:r1], transB=True, b_weight=True)
                    ops.gemm(oc, dfc[r0:r1], self._fc_kernel.grad, transA=True, beta=1.0)
                    ops.colsum(dfc[r0:r1], self._fc_bias.grad, accumulate=True)
                finally:
                    ops.set_sm_budget(0)
                ev = rnn._event(f'fwd hook steps {t0}..{t1}')
                ev.record(bulk)
                for c in range(t0 // C, (t1 + C - 1) // C):
                    dout_ready.setdefault(c, []).append(ev)
                if is_last:
                    last[0] = ev

        outs, rnn_state = rnn.forward_sequence(inputs.contiguous(), keep=keep, u=u_drop, seed=seed, chunk_hook=hook)
        self._outs = outs
        self._state_from_fc(fc, rnn_state)
        if rnn.PIPE_DESCENDING and rnn.use_pipeline(T, B):
            # BPTT of a chunk waits for that chunk's dout only; the hooks still in flight run beside it on the bulk stream,
            # and backward_sequence ends with main waiting for the bulk stream
            dx = rnn.backward_sequence(dout.view(T, B, -1), need_dx=need_dx, pipelined=True, dout_ready=dout_ready)
            main.wait_event(last[0])
            ops.sum_into(nll, ws['loss'], scale=gscale)
        else:
            main.wait_event(last[0])
            ops.sum_into(nll, ws['loss'], scale=gscale)
            dx = rnn.backward_sequence(dout.view(T, B, -1), need_dx=need_dx, pipelined=True)
        return ws['loss'], nll, dx

    # -------------------------------------------------------------- generation
    def generate(self, x, num_steps, u=None, seed=0):
        """rnn_estimator.py:271-323. The intro scan runs through the sequence kernels; the num_steps-step scan of
        {sample_single, single_step} then runs as ONE launch of the persistent kernel `mnn_generate_fused` (samplers, LSTM
        layers and the Dense layer as CTA groups with resident weights) when the batch fits one 128-row slab and the sampled
        frame is the layer-0 input (Composer, Jamming, Joint-with-NADE; not the Feedback modes, whose generators take
        [sample ; feedback]); otherwise, and in parity runs with supplied uniforms unless ops.GENERATE_MODE == 'fused', as the
        per-step loop of the base class (8 launches per step, fp32-accurate)."""
        M, D, H = self._num_tracks, self._num_dims, self._num_hidden[-1]
        B = x.shape[1]
        rnn = self._rnn
        fits = ops.generate_fused_supported(rnn.num_layers, self._num_inputs, rnn.num_units, B, M, D, H)
        mode = ops.GENERATE_MODE
        if not fits or mode == 'steps' or (mode == 'auto' and u is not None):
            return super().generate(x, num_steps, u=u, seed=seed)
        state = self._get_state(x, lengths=None, last_outputs=True)
        fc = self._fc_of(state).contiguous()
        out = torch.empty(B, num_steps, M * D, device=x.device)
        ops.generate_fused([k.data for k in rnn.kernels], [b.data for b in rnn.biases],
                           [(s.c.contiguous(), s.h.contiguous()) for s in state.rnn_state], self._fc_kernel.data,
                           self._fc_bias.data, self._bank.w_enc.data, self._bank.w_dec.data, fc, out,
                           u=None if u is None else u.contiguous(), use_philox=True, seed=seed, offset0=0)
        return out

    def single_step(self, inputs, initial_state):
        """rnn_nade.py:253-277: one RNN step + Dense -> new biases."""
        out, rnn_state = self._rnn.step(inputs, initial_state.rnn_state)
        fc = torch.empty(inputs.shape[0], self._fc_kernel.shape[1], device=inputs.device)
        ops.gemm(out, self._fc_kernel.data, fc, bias=self._fc_bias.data)
        return self._state_from_fc(fc, rnn_state)

    def sample_single(self, inputs, state, u=None, seed=0, offset=0, out=None, temperature=1.0):
        """rnn_nade.py:304-318 / rnn_multinade.py:292-317: NADE.sample(temperature=1.) per track, tracks stacked
        with axis=2 then flattened: feature index d*M + m. u[M,B,D] uniforms or None (Philox). `inputs` unused."""
        fc = self._fc_of(state)
        B = fc.shape[0]
        M, D = self._num_tracks, self._num_dims
        if out is None:
            out = torch.empty(B, M * D, device=fc.device)
        nll = torch.empty(M, B, device=fc.device)
        sampling = temperature is not None
        ops.nade_sample(fc, self.enc_col0, self.dec_col0, self._bank.w_enc.data, self._bank.w_dec.data, out,
                        out.stride(0), M, 1, u=u.contiguous() if (sampling and u is not None) else None,
                        use_philox=sampling and u is None, seed=seed, offset=offset, nll=nll)
        return out, nll
