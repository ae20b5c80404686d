"""MultINN core (mirrors reference models/multinn/core/{multinn_interface,multinn_core,multi_encoder_nn}.py).

The reference builds one TF graph around placeholders x[B,T,D,M], lengths[B], is_train and runs it with
sess.run; here the same pipeline (inputs -> encoders -> generators -> metrics, multinn_core.py:178-244) is a
set of eager methods on device tensors:
  train_generators(optimizer, lr) -> step(x, ...) -> loss      (one sess.run of train.py:186-189)
  evaluate(x) -> {'batch/loss', 'log_likelihood', 'nll'}       (is_train=False forward)
  generate(x_intro, num_steps) / sampler(num_beats)            (multinn_core.py:324-341)
Inputs are staged once per call by the K0 kernel (zero-pad, unstack/stack, shift, target bit masks).
"""
import abc
import os

import torch

from .. import ops
from ..common.model import Model
from ..encoders.pass_encoder import PassEncoder
from ..params import ParamArena
from ..training import AdamOptimizer, GradientApplier, GradientDescentOptimizer, dp_row_map


class MultINNCore(Model, abc.ABC):
    def __init__(self, config, params, name='MultINN', device='cuda', seed=23):
        super().__init__(name=name)
        self._mode = 'core'
        self._config, self._params = config, params
        self._device = torch.device(device)
        self._encoder_type = params['encoder']['type']
        self._generator_type = params['generator']['type']
        if self._encoder_type == 'Pass':
            encoder_class = PassEncoder
        elif self._encoder_type in ('RBM', 'DBN'):
            from ..encoders.dbn_encoder import DBNEncoder
            encoder_class = DBNEncoder
        else:
            raise ValueError('Incorrect encoder type, supported types are `Pass`, `RBM`, and `DBN`')
        if self._generator_type == 'RBM':
            generator_class = 'RBM'          # resolved lazily by the modes that support it (joint, jamming, ...)
        elif self._generator_type == 'NADE':
            from ..generators.rnn_nade import RnnNade
            generator_class = RnnNade
        else:
            raise ValueError('Incorrect generator type, supported types are `RBM`, and `NADE`')
        pr = config['data']['pitch_range']
        self._num_dims = (pr['highest'] - pr['lowest']) * config['training']['num_pixels']   # multinn_core.py:57-58
        self._tracks = list(config['data']['instruments'])
        self._feedback_module = False
        self._keep_prob = params['keep_prob']
        self._tune_encoder = params['tune_encoder']
        self._placeholders = {'x': None, 'lengths': None, 'is_train': None}
        self._enc_arena = ParamArena()      # encoders are trained separately (train_encoders.py) and frozen here
        self._arena = ParamArena()          # generators (+ feedback module): one flat bucket
        self._encoders = self._init_encoders(encoder_class)
        self._generators = self._init_generators(generator_class)
        self._enc_arena.finalize(self._device, seed=seed + 1)
        self._arena.finalize(self._device, seed=seed)
        self._applier = None
        self._stage = {}

    # ------------------------------------------------------------------ properties (multinn_interface.py)
    mode = property(lambda s: s._mode)
    num_dims = property(lambda s: s._num_dims)
    tracks = property(lambda s: s._tracks)
    num_tracks = property(lambda s: len(s._tracks))
    encoder_type = property(lambda s: s._encoder_type)
    generator_type = property(lambda s: s._generator_type)
    encoders = property(lambda s: s._encoders)
    generators = property(lambda s: s._generators)
    feedback_module = property(lambda s: s._feedback_module)
    keep_prob = property(lambda s: s._keep_prob)
    tune_encoder = property(lambda s: s._tune_encoder)
    placeholders = property(lambda s: s._placeholders)
    arena = property(lambda s: s._arena)
    encoder_arena = property(lambda s: s._enc_arena)

    @property
    def trainable_params(self):
        return list(self._arena.params)

    @abc.abstractmethod
    def _init_encoders(self, encoder_class):
        ...

    @abc.abstractmethod
    def _init_generators(self, generator_class):
        ...

    _supports_lengths = False
    TRAIN_GEMM_SPLIT = os.environ.get('MNN_TRAIN_GEMM_SPLIT', 'pair')

    # ------------------------------------------------------------------ input staging (K0)
    def _check_x(self, x, lengths):
        if x.dim() != 4 or x.shape[2] != self.num_dims or x.shape[3] != self.num_tracks:
            raise ValueError(f'x must be [batch, time, {self.num_dims}, {self.num_tracks}], got {tuple(x.shape)}')
        if lengths is not None and int(torch.as_tensor(lengths).min()) < x.shape[1] and not self._supports_lengths:
            raise NotImplementedError(f'variable sequence lengths are not supported in {self._mode} mode yet '
                                      '(the NADE generators of the Composer, Jamming and Feedback modes support them)')
        if not x.is_cuda:
            raise ValueError('x must be a CUDA tensor: multinn_b200 has no CPU path')
        if x.dtype in (torch.uint8, torch.bool):          # bool / byte piano-rolls as stored by prepare_data.py:56
            return x.contiguous()
        return x.contiguous().float()

    def _stage_inputs(self, x, stacked=False, per_track=False, bits=False):
        """core/multi_encoder_nn.py:66-87: zero-pad one step in front of time + unstack tracks (+ Composer's
        stack/reshape, multinn_composer.py:73-80). Returns time-major tensors (T+1 slots)."""
        B, T, D, M = x.shape
        key = (B, T)
        st = self._stage.get(key)
        if st is None:                        # one shape at a time: a new (B, T) releases the previous staging buffers
            st = {}
            self._stage = {key: st}
        dev = x.device
        if stacked and 'xin' not in st:
            st['xin'] = torch.empty(T + 1, B, D * M, device=dev)
        if per_track and 'xtr' not in st:
            st['xtr'] = torch.empty(M, T + 1, B, D, device=dev)
        if bits and 'bits' not in st:
            st['bits'] = torch.empty(M, T * B, 4, dtype=torch.int32, device=dev)
        ops.pack_pianoroll(x, st['xin'] if stacked else None, st['xtr'] if per_track else None,
                           st['bits'] if bits else None)
        if stacked and ops.BF16_INPUT_TWIN and ops._gemm_split == 'pair':
            # training step: the binary stacked rows once more as an exact bf16 plane -- the layer-0 projection and the
            # x-rows weight-gradient GEMM read it through TMA without the in-kernel operand split (ops.gemm, a_exact)
            if 'xin16' not in st:
                st['xin16'] = torch.empty((T + 1) * B, (D * M + 7) // 8 * 8, dtype=torch.int16, device=dev)
            ops.pack_stacked_bf16(x, st['xin16'])
            ops.register_twin(st['xin'].view((T + 1) * B, D * M), st['xin16'])
        return st

    # The per-track generators of Jamming / Feedback(-RNN) are independent between the encode and the feedback backward
    # (multinn_jamming.py:60-68, multinn_feedback.py:86-94 build M separate sub-graphs that TF may run concurrently): each
    # runs on its own stream. At the data-parallel shard sizes a generator's recurrences are latency chains on 16-64 SMs,
    # so five of them side by side cost little more than one (C4 on 8 GPUs: 51 -> see DESIGN.md section 10).
    TRACK_STREAMS = os.environ.get('MNN_TRACK_STREAMS', '1') != '0'

    def _per_track(self, fn):
        """[fn(m, generator) for every generator], each call on its own stream between two joins with the current one."""
        gens = self._generators
        if not self.TRACK_STREAMS or len(gens) < 2 or not torch.cuda.is_available():
            return [fn(m, g) for m, g in enumerate(gens)]
        main = torch.cuda.current_stream()
        streams = self.__dict__.setdefault('_track_streams', [torch.cuda.Stream() for _ in gens])
        start = torch.cuda.Event()
        start.record(main)
        out = []
        for m, (g, st) in enumerate(zip(gens, streams)):
            with torch.cuda.stream(st):
                st.wait_event(start)
                out.append(fn(m, g))
        for st in streams:
            main.wait_stream(st)
        return out

    def _encode_tracks(self, x, u_enc=None, seed=0, need_tracks=True, need_stack=True):
        """core/multi_encoder_nn.py:66-115: zero-pad, unstack, encode every track with its own encoder (PassEncoder:
        identity; DBNEncoder: SAMPLED last-layer codes, stop_gradient unless tune_encoder). Returns
        (xe[M][(T+1),B,E] per-track encodings, stack[(T+1),B,E*M] with feature e*M + m as multinn_composer.py:73-80 /
        multinn_feedback.py:67-73 stack them, bits[M,T*B,4] target masks of xe[m][1:])."""
        B, T, D, M = x.shape
        if self.encoder_type == 'Pass':   # identity: stage only the layouts the mode reads (each is a full pass over HBM)
            st = self._stage_inputs(x, stacked=need_stack, per_track=need_tracks, bits=True)
            return ([st['xtr'][m] for m in range(M)] if need_tracks else None, st['xin'] if need_stack else None,
                    st['bits'])
        st = self._stage_inputs(x, per_track=True)
        xe = []
        for m, enc in enumerate(self._encoders):
            _, h = enc.encode(st['xtr'][m].view((T + 1) * B, D), u=None if u_enc is None else u_enc[m], seed=seed + 31 * m)
            xe.append(h.view(T + 1, B, -1))
        stack = torch.stack(xe, dim=3).reshape(T + 1, B, -1)
        bits = torch.empty(M, T * B, 4, dtype=torch.int32, device=x.device)
        for m in range(M):
            ops.pack_rows(xe[m][1:].reshape(T * B, -1), bits[m])
        return xe, stack, bits

    def _decode_tracks(self, samples_h, u_dec=None, seed=0):
        """Per-track `encoder.decode` of generated codes samples_h[B,S,E,M] -> music[B,S,D,M] (multinn_composer.py:140-150,
        multinn_jamming.py:125-132, multinn_feedback.py:166-172); identity for Pass encoders."""
        B, S, E, M = samples_h.shape
        if self.encoder_type == 'Pass':
            return samples_h
        music = torch.empty(B, S, self.num_dims, M, device=samples_h.device)
        with ops.row_map_scaled(S):                        # decode rows are b-major (b*S + s)
            for m, enc in enumerate(self._encoders):
                h_m = samples_h[..., m].reshape(B * S, E).contiguous()      # the track slice is a strided view
                _, v = enc.decode(h_m, u=None if u_dec is None else u_dec[m], seed=seed + 977 * m)
                music[..., m] = v.view(B, S, self.num_dims)
        return music

    @staticmethod
    def rows_to_reference_order(t, T, B):
        """[..., T*B] time-major rows -> [B*T, ...] in the reference's flatten order n = b*T + t
        (utils/sequences.py:22-24), moving the leading (track) axis last."""
        M = t.shape[0]
        return t.view(M, T, B).permute(2, 1, 0).reshape(B * T, M)

    @staticmethod
    def valid_rows(lengths, T, B, device):
        """Indices of the rows flatten_maybe_padded_sequences keeps, in its order n = b*T + t (utils/sequences.py:29-31);
        None for full lengths."""
        if lengths is None:
            return None
        lengths = torch.as_tensor(lengths).to('cpu', torch.int64)
        if int(lengths.min()) >= T:
            return None
        mask = torch.arange(T)[None, :] < lengths[:, None]
        return mask.reshape(-1).nonzero().squeeze(1).to(device)

    # ------------------------------------------------------------------ training
    def _make_optimizer(self, optimizer, lr):
        if isinstance(optimizer, str):
            optimizer = AdamOptimizer(lr) if optimizer.lower() == 'adam' else GradientDescentOptimizer(lr)
        return optimizer

    def train_generators(self, optimizer='adam', lr=0.01, separate_losses=False):
        """Returns `step(x, lengths=None, u_drop=None, seed=None) -> loss` (device scalar): one fwd + bwd +
        allreduce + clip + apply, the unit train.py:186-189 runs per batch piece. Quirk Q7: always joint."""
        opt = self._make_optimizer(optimizer, lr)
        self._applier = GradientApplier(self._arena, opt, lr=lr)
        counter = [0]

        def step(x, lengths=None, u_drop=None, seed=None, keep=None, row_base=None, global_batch=None, **extra):
            x = self._check_x(x, lengths)
            self._applier.zero_grad()
            s = counter[0] if seed is None else seed      # the same seed on every rank: noise is keyed by the global row
            counter[0] += 1
            if lengths is not None and self._supports_lengths:
                extra['lengths'] = lengths
            # training GEMMs on bf16 pairs (~2^-17 per product, 9 % faster; parity bar 1e-4); evaluate() / generate() keep
            # the 2.5-product split. TRAIN_GEMM_SPLIT = '2.5' switches it off.
            with ops.row_map(*dp_row_map(x.shape[0], row_base, global_batch)), ops.gemm_split(self.TRAIN_GEMM_SPLIT):
                loss = self._forward_backward(x, keep=self._keep_prob if keep is None else keep, u_drop=u_drop,
                                              seed=s * 1000003, **extra)
            self._applier.apply()
            self._metrics['batch/loss'] = loss
            return loss

        return step

    def _pretrain_rows(self, x, seed):
        """(generator, flattened input rows) pairs for pretrain_generators; None rows for generators without a module to
        pre-train (every NADE generator)."""
        return [(g, None) for g in self._generators]

    def pretrain_generators(self, optimizer=None, lr=0.01, separate_losses=False):
        """multinn.py:232-250, multinn_composer.py:153-171, multinn_jamming.py:135-154, multinn_joint.py: returns
        `step(x, lengths=None, u=None, seed=None) -> number of generators updated`. RNN-NADE generators have nothing to
        pre-train (rnn_nade.py:320-326: empty update ops), an RNN-RBM generator gets one CD-k update of its RBM on the
        flattened input frames (rnn_rbm.py:299-322; `RBM.train` applies it itself, the optimizer is unused there too)."""
        counter = [0]

        def step(x, lengths=None, u=None, seed=None, row_base=None, global_batch=None):
            x = self._check_x(x, lengths)
            s = counter[0] if seed is None else seed
            counter[0] += 1
            updated = 0
            with ops.row_map(*dp_row_map(x.shape[0], row_base, global_batch)):
                for gen, rows in self._pretrain_rows(x, s * 1000003):
                    if rows is not None and gen.pretrain(rows, lr, u=u, seed=s * 7919 + 13) is not None:
                        updated += 1
            return updated

        return step

    # ------------------------------------------------------------------ encoder pre-training (train_encoders.py)
    def _encoder_rows(self, x, lengths=None):
        """Per-encoder training rows: the zero-padded inputs [B,T+1,..] flattened by flatten_maybe_padded_sequences
        with the UNPADDED lengths (encoder.build(x=self._inputs[i], lengths), dbn_encoder.py:212-214): the rows
        t < lengths[b] of the padded sequence, i.e. the zero frame and the first lengths[b] - 1 real frames. Returns a
        list with one [rows, num_dims] tensor per encoder (CD sums do not depend on the row order)."""
        B, T, D, M = x.shape
        keep = None
        if lengths is not None and int(torch.as_tensor(lengths).min()) < T + 1:
            ln = torch.as_tensor(lengths).to('cpu', torch.int64)
            mask = (torch.arange(T + 1)[:, None] < ln[None, :]).reshape(-1)          # time-major rows t*B + b
            keep = mask.nonzero().squeeze(1).to(x.device)
        if len(self._encoders) == 1:                                                 # Joint: one encoder over D*M dims
            st = self._stage_inputs(x, stacked=True)
            rows = [st['xin'].view((T + 1) * B, D * M)]
        else:
            st = self._stage_inputs(x, per_track=True)
            rows = [st['xtr'][m].view((T + 1) * B, D) for m in range(M)]
        if keep is None:
            return [r[:T * B] for r in rows]                                         # full lengths: t < T
        return [r.index_select(0, keep) for r in rows]

    def train_encoders(self, optimizer=None, lr=0.01, layer=0):
        """core/multi_encoder_nn.py:155-195 / multinn_joint.py: the encoders are pre-trained separately but in parallel,
        layer by layer, with CD-k (`RBM.train` applies the update itself: the optimizer argument is unused there too).
        Returns (init, step): `init(x, lengths=None)` = the init_ops (visible-bias initialisation, run once on
        X_train[:1600], train_encoders.py:106-109); `step(x, lengths=None, u=None, seed=None)` = one CD update of
        `layer` on every encoder, returning {'batch/loss', 'log_likelihood'} averaged over the encoders
        (`_combine_track_metrics`). Pass encoders (no parameters) make both no-ops."""
        trainable = [e for e in self._encoders if getattr(e, 'stochastic', False)]
        counter = [0]

        def init(x, lengths=None, u=None, seed=0):
            x = self._check_encoder_x(x)
            for m, (enc, rows) in enumerate(zip(self._encoders, self._encoder_rows(x, lengths))):
                if enc in trainable:
                    enc.init_bias(rows, layer, u=None if u is None else u[m], seed=seed + 31 * m)

        def step(x, lengths=None, u=None, seed=None):
            x = self._check_encoder_x(x)
            s = counter[0] if seed is None else seed
            counter[0] += 1
            out = {'batch/loss': torch.zeros((), device=x.device), 'log_likelihood': torch.zeros((), device=x.device)}
            for m, (enc, rows) in enumerate(zip(self._encoders, self._encoder_rows(x, lengths))):
                if enc not in trainable:
                    continue
                um = None if u is None else u[m]
                with ops.row_map(*dp_row_map(x.shape[0])):
                    met = enc.layer_metrics(rows, layer, u=None if um is None else um['metrics'], seed=s * 7919 + 31 * m)
                    enc.train(rows, lr, layer=layer, u=None if um is None else um['train'], seed=s * 7919 + 31 * m + 1)
                for k in out:
                    out[k] = out[k] + met[k].reshape(()) / max(len(trainable), 1)
            self._metrics.update({f'encoders/{k}': v for k, v in out.items()})
            return out

        return init, step

    def evaluate_encoders(self, x, lengths=None, layer=0, u=None, seed=0):
        """Encoder metrics without an update (the metrics_upd ops collect_metrics runs in train_encoders.py:178-190):
        {'batch/loss', 'log_likelihood'} of RBM `layer`, averaged over the encoders, plus the number of rows."""
        x = self._check_encoder_x(x)
        trainable = [e for e in self._encoders if getattr(e, 'stochastic', False)]
        out = {'batch/loss': torch.zeros((), device=x.device), 'log_likelihood': torch.zeros((), device=x.device)}
        rows_n = 0
        for m, (enc, rows) in enumerate(zip(self._encoders, self._encoder_rows(x, lengths))):
            if enc not in trainable:
                continue
            met = enc.layer_metrics(rows, layer, u=None if u is None else u[m], seed=seed + 31 * m)
            rows_n = rows.shape[0]
            for k in out:
                out[k] = out[k] + met[k].reshape(()) / len(trainable)
        out['rows'] = rows_n
        return out

    def _check_encoder_x(self, x):
        if x.dim() != 4 or x.shape[2] != self.num_dims or x.shape[3] != self.num_tracks:
            raise ValueError(f'x must be [batch, time, {self.num_dims}, {self.num_tracks}], got {tuple(x.shape)}')
        if not x.is_cuda:
            raise ValueError('x must be a CUDA tensor: multinn_b200 has no CPU path')
        return x.contiguous() if x.dtype in (torch.uint8, torch.bool) else x.contiguous().float()

    @abc.abstractmethod
    def _forward_backward(self, x, keep, u_drop, seed, **extra):
        """One fwd+bwd over the batch; `extra` carries mode-specific uniform-noise tensors for parity runs."""

    @abc.abstractmethod
    def evaluate(self, x, lengths=None):
        ...

    @abc.abstractmethod
    def generate(self, x, num_steps, u=None, seed=0):
        ...

    def sampler(self, num_beats):
        """multinn_core.py:324-341: returns `sample(x_intro, u=None) -> [B,num_steps,D,M]`."""
        dc = self._config['data']
        pitch_span = dc['pitch_range']['highest'] - dc['pitch_range']['lowest']
        num_steps = num_beats * dc['beat_resolution'] * pitch_span // self.num_dims
        def sample(x, u=None, seed=None, row_base=None, global_batch=None):
            with ops.row_map(*dp_row_map(x.shape[0], row_base, global_batch)):
                return self.generate(x, num_steps, u=u, seed=0 if seed is None else seed)
        return sample

    def evaluator(self):
        """multinn_core.py:343-362: returns `evaluate_music(x) -> {summary scope: value}`: x[B,T,num_dims,M] (device tensor
        or array, data or generated samples) reshaped into bar music and scored with the musical metrics
        (metrics/musical_tf.py:186-229). Reporting code: runs on the host in NumPy, once per sampling run."""
        from ..metrics import musical
        dc = self._config['data']
        pitch_span = dc['pitch_range']['highest'] - dc['pitch_range']['lowest']
        beat_resolution, tracks = dc['beat_resolution'], self.tracks

        def evaluate_music(x):
            x = x.detach().cpu().numpy() if torch.is_tensor(x) else x
            return musical.metric_summary(musical.to_bars(x, beat_resolution, pitch_span), tracks)

        return evaluate_music

    # ------------------------------------------------------------------ checkpoints (model.py:180-234: trainable vars only)
    def state_dict(self):
        return {'generators': self._arena.state_dict(), 'encoders': self._enc_arena.state_dict()}

    def load_state_dict(self, sd):
        self._arena.load_state_dict(sd.get('generators', {}))
        self._enc_arena.load_state_dict(sd.get('encoders', {}))

    def save(self, path):
        torch.save(self.state_dict(), path)

    def load(self, path):
        self.load_state_dict(torch.load(path, map_location='cpu'))

    def load_tf(self, path, which='generators', name_map=None, strict=False):
        """Restores the generator (or encoder) parameters from a TensorFlow Saver checkpoint of the reference
        (common/model.py:216-234, core/multinn_core.py:425-448: `path` = its directory or prefix), read in pure Python.
        strict=False by default: a Saver file may also hold optimiser slots and counters. Returns the applied mapping."""
        from ..utils import tf_checkpoint, tf_import
        return tf_import.load_tf_variables(self, tf_checkpoint.read_checkpoint(path), which=which, name_map=name_map,
                                           strict=strict)

    def save_tf(self, prefix, which='generators'):
        """Writes the generator (or encoder) parameters as a TF V2 checkpoint (`<prefix>.index`, `.data-00000-of-00001`,
        `checkpoint`) with TF-shaped tensors, the files common/model.py:180-214 writes."""
        from ..utils import tf_checkpoint, tf_import
        tf_checkpoint.write_checkpoint(prefix, tf_import.export_tf_variables(self, which))
