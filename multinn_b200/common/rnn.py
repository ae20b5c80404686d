"""Temporal unit: stacked LSTM with output dropout (mirrors reference models/common/rnn.py:13-219).

Cell semantics = CudnnCompatibleLSTMCell == LSTMBlockCell(forget_bias=0): kernel[(I+R),4R] with rows
[x ; h] and gate column blocks i, j, f, o; DropoutWrapper(output_keep_prob) on every layer's output
(state h is not dropped); MultiRNNCell stacking (rnn.py:104-145).

B200 layout: sequences are time-major [T,B,*] so that one step's rows are contiguous; the input projection
x.Wx + b of a whole sequence is ONE GEMM over all T*B rows (hoisted out of the recurrence), only h.Wh is
sequential. Gate pre-activations are overwritten by their activations (saved for BPTT) and then by their
gradients, so the recurrence keeps a single [T,B,4R] buffer per layer.
"""
import os

import torch

from .. import ops
from ..params import glorot_uniform, zeros


class LSTMStateTuple(tuple):
    """(c, h) like tf.nn.rnn_cell.LSTMStateTuple."""
    __slots__ = ()

    def __new__(cls, c, h):
        return tuple.__new__(cls, (c, h))

    c = property(lambda s: s[0])
    h = property(lambda s: s[1])


class RNN:
    """`RNN(num_units, keep_prob, name)`; `num_units` is the list of layer sizes (rnn.py:27-60)."""

    def __init__(self, arena, num_inputs, num_units, keep_prob=1.0, name='rnn', binary_inputs=False):
        if isinstance(num_units, int):
            num_units = [num_units]
        self.name = name
        self._num_units = list(num_units)
        self._num_inputs = num_inputs
        self._keep_prob = keep_prob
        self._binary_inputs = binary_inputs     # layer-0 inputs are exactly {0,1}: 2 tf32 products instead of 3
        self.kernels, self.biases = [], []
        i = num_inputs
        for l, r in enumerate(self._num_units):
            # TF variable names: .../multi_rnn_cell/cell_l/cudnn_compatible_lstm_cell/{kernel,bias}
            self.kernels.append(arena.add(f'{name}/cell_{l}/kernel', (i + r, 4 * r), glorot_uniform(i + r, 4 * r)))
            self.biases.append(arena.add(f'{name}/cell_{l}/bias', (4 * r,), zeros()))
            i = r
        self._ws = {}
        self._saved = None

    @property
    def num_units(self):
        return self._num_units

    @property
    def num_layers(self):
        return len(self._num_units)

    @property
    def keep_prob(self):
        return self._keep_prob

    @property
    def trainable_params(self):
        return [p for pair in zip(self.kernels, self.biases) for p in pair]

    def in_dims(self):
        return [self._num_inputs] + self._num_units[:-1]

    def zero_state(self, batch_size, device):
        """rnn.py:155-176 (learn_zero_state=False): zeros for every layer."""
        return [LSTMStateTuple(torch.zeros(batch_size, r, device=device), torch.zeros(batch_size, r, device=device))
                for r in self._num_units]

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, T, B, device, dropout):
        key = (T, B, bool(dropout))
        ws = self._ws.get(key)
        if ws is None:
            ws = []
            for r in self._num_units:
                d = dict(gates=torch.empty(T, B, 4 * r, device=device),
                         hbuf=torch.empty(T + 1, B, r, device=device),
                         cbuf=torch.empty(T + 1, B, r, device=device))
                if dropout:
                    d['out'] = torch.empty(T, B, r, device=device)
                    d['dscale'] = torch.empty(T, B, r, device=device)
                ws.append(d)
            self._ws = {key: ws}  # keep one shape resident
        return ws

    # ------------------------------------------------------------------ whole-sequence forward / BPTT
    # At small per-GPU batches (data-parallel shards) a recurrence step is a fixed latency chain that keeps well under
    # half of the SMs busy, so the layers are run as a WAVEFRONT over time chunks: layer l works on chunk c while
    # layer l+1 works on chunk c-1 on another stream (SURVEY 7 "hard parts": recurrence latency at B/8).
    WAVEFRONT_MAX_BATCH = int(os.environ.get('MNN_WAVEFRONT_MAX_BATCH', 1024))   # forward; measured gain up to 1024
    WAVEFRONT_MAX_BATCH_BWD = 512                                                # BPTT: the SM budgets cost more above
    WAVEFRONT_CHUNK = int(os.environ.get('MNN_WAVEFRONT_CHUNK', 32))
    WAVEFRONT_GEMM_SMS = 32

    # Time-chunk PIPELINE (small per-GPU batches): while the recurrences walk the chunks on their (high-priority)
    # streams, the batched work either side of them -- Dense, NADE forward/backward, data- and weight-gradient GEMMs,
    # bias column sums -- runs chunk by chunk on a low-priority "bulk" stream inside an SM budget that leaves the
    # recurrence kernels' SMs free. The recurrences hold ~100 SMs but are a latency chain; the bulk work fills the rest.
    PIPE_MAX_BATCH = int(os.environ.get('MNN_PIPE_MAX_BATCH', 512))
    PIPE_TAIL_STEPS = int(os.environ.get('MNN_PIPE_TAIL_STEPS', 16))     # large batches: steps whose bulk work runs beside the top layer
    PIPE_MAX_BATCH_FWD = int(os.environ.get('MNN_PIPE_MAX_BATCH_FWD', 1024))
    PIPE_FWD_BUDGETS = tuple(int(v) for v in os.environ.get('MNN_PIPE_FWD_BUDGETS', '96,48').split(','))
    PIPE_BWD_BUDGETS = tuple(int(v) for v in os.environ.get('MNN_PIPE_BWD_BUDGETS', '0').split(','))   # 0: by batch
    PIPE_BULK_SMS_FWD = int(os.environ.get('MNN_PIPE_BULK_SMS_FWD', 0))    # 0: every SM the recurrence kernels leave
    PIPE_BULK_SMS_BWD = int(os.environ.get('MNN_PIPE_BULK_SMS_BWD', 0))
    PIPE_AUX_SMS = int(os.environ.get('MNN_PIPE_AUX_SMS', 16))          # dx GEMMs between the BPTT layers (own stream); 0: off
    PIPE_SLOW_HOOKS = int(os.environ.get('MNN_PIPE_SLOW_HOOKS', 0))     # forward chunk hooks under the recurrence; 0: by batch
    PIPE_FULL_WGRADS = int(os.environ.get('MNN_PIPE_FULL_WGRADS', 3))   # last BPTT chunks whose weight grads use every SM
    # after the forward recurrences: consumer work of the remaining chunks in DESCENDING chunk order with one event per
    # chunk, so that BPTT (t descending) starts on the last chunk while the earlier chunks' work runs beside it.
    # Measured on a B200 (gpurun_out/r2n, B=256): 21.9 ms per step against 14.0 ms for one full-width call before BPTT --
    # the bulk work is throughput-bound and loses more to chunking under an SM budget than the overlap wins. Off.
    PIPE_DESCENDING = os.environ.get('MNN_PIPE_DESCENDING', '0') != '0'

    COLSUM_SIDE_STREAM = os.environ.get('MNN_COLSUM_SIDE', '1') != '0'

    TRACE = None        # debug (tools/pipeline_trace.py): list collecting (label, timing event) of the chunk schedule

    @classmethod
    def _event(cls, label=None):
        if cls.TRACE is None:
            return torch.cuda.Event()
        ev = torch.cuda.Event(enable_timing=True)
        cls.TRACE.append((label, ev))
        return ev

    def _use_wavefront(self, T, B, backward=False):
        mb = min(self.WAVEFRONT_MAX_BATCH, self.WAVEFRONT_MAX_BATCH_BWD) if backward else self.WAVEFRONT_MAX_BATCH
        return (self.num_layers > 1 and B <= mb and T % self.WAVEFRONT_CHUNK == 0 and T >= 2 * self.WAVEFRONT_CHUNK)

    def use_pipeline(self, T, B):
        """Backward half of the pipeline (chunked BPTT wavefront with the weight gradients beside it)."""
        return B <= self.PIPE_MAX_BATCH and self._use_wavefront(T, B) and self._use_wavefront(T, B, backward=True)

    def use_any_pipeline(self, T, B):
        """Does `forward_sequence` call a chunk hook for this shape (forward wavefront pipeline or top-layer tail)?"""
        if self.use_fwd_pipeline(T, B):
            return True
        return (self.PIPE_TAIL_STEPS > 0 and self.PIPE_MAX_BATCH > 0 and not self._use_wavefront(T, B)
                and T % self.WAVEFRONT_CHUNK == 0 and T >= 2 * self.WAVEFRONT_CHUNK and B >= 1024)

    def use_fwd_pipeline(self, T, B):
        """Forward half (chunk hooks beside the forward wavefront): also where BPTT runs layer by layer (B = 1024)."""
        return self.PIPE_MAX_BATCH > 0 and B <= max(self.PIPE_MAX_BATCH, self.PIPE_MAX_BATCH_FWD) and self._use_wavefront(T, B)

    def _streams(self, device):
        st = self.__dict__.get('_wave_streams')
        if st is None:
            st = self._wave_streams = [torch.cuda.Stream(device=device, priority=-1) for _ in range(self.num_layers)]
        return st

    def _bwd_budgets(self, B):
        """SM budgets of the two BPTT kernels in pipeline mode. Measured: at B=256 (64, 32) gives the same K-splits as
        (72, 36) and frees 12 SMs for the bulk stream; at B=512 (48, 24) beats both."""
        if self.PIPE_BWD_BUDGETS[0] > 0:
            return self.PIPE_BWD_BUDGETS
        return (64, 32) if B < 512 else (48, 24)

    def bulk_budget(self, T, B, backward=False):
        """SM budget of the bulk stream while the recurrences run: what the layers' persistent kernels (and the budgeted
        GEMM that shares a layer's stream) leave free."""
        fixed = self.PIPE_BULK_SMS_BWD if backward else self.PIPE_BULK_SMS_FWD
        if fixed > 0:
            return fixed
        key = (T, B, backward)
        cache = self.__dict__.setdefault('_bulk_budgets', {})
        if key not in cache:
            budgets = self._bwd_budgets(B) if backward else self.PIPE_FWD_BUDGETS
            used = 0
            for l, r in enumerate(self._num_units):
                n = ops.lstm_seq_ctas(self.WAVEFRONT_CHUNK, B, r, budgets[min(l, len(budgets) - 1)], backward)
                if backward and self.PIPE_AUX_SMS > 0:
                    used += n + (self.PIPE_AUX_SMS if l == 1 else 0)   # dx GEMMs run on their own stream
                else:                                                  # input projection / dx GEMM on the layer's stream
                    used += max(n, self.WAVEFRONT_GEMM_SMS) if l > 0 else n
            cache[key] = max(16, ops.num_sms() - used)
        return cache[key]

    def aux_stream(self, device):
        st = self.__dict__.get('_aux_stream')
        if st is None:
            st = self._aux_stream = torch.cuda.Stream(device=device, priority=-1)
        return st

    def bulk_stream(self, device):
        st = self.__dict__.get('_bulk_stream')
        if st is None:
            st = self._bulk_stream = torch.cuda.Stream(device=device, priority=0)
        return st

    def forward_sequence(self, x, keep=1.0, u=None, seed=0, initial_state=None, chunk_hook=None):
        """x[T,B,I] time-major -> outputs[T,B,R_top] (dropped out when keep < 1), final state.
        u: optional list (per layer) of [T,B,R_l] uniforms for reproducible dropout; else Philox(seed).
        Saves what BPTT needs (call `backward_sequence` next).
        chunk_hook(t0, t1, done_event, outs_top, budget, last): pipeline modes (`use_any_pipeline`); called after the top layer
        has been enqueued up to step t1, the hook enqueues the consumer work of steps [t0, t1) on `bulk_stream` behind
        `done_event` under the SM `budget` (0 = every SM). The first PIPE_SLOW_HOOKS chunks get one budgeted call each
        (they run beside the recurrences), the rest ONE call at full width once the recurrences are through."""
        T, B, I = x.shape
        assert I == self._num_inputs and x.is_contiguous()
        dropout = keep < 1.0
        ws = self._workspace(T, B, x.device, dropout)
        for l, w in enumerate(ws):
            if initial_state is None:
                w['hbuf'][0].zero_()
                w['cbuf'][0].zero_()
            else:
                w['cbuf'][0].copy_(initial_state[l][0])
                w['hbuf'][0].copy_(initial_state[l][1])
        outs = [w['out'] if dropout else w['hbuf'][1:] for w in ws]
        # layer 0: hoisted input projection over all T*B rows (pipeline mode: chunk 0 now, the rest beside the recurrences)
        r0 = self._num_units[0]
        hook = chunk_hook if (chunk_hook is not None and self.use_fwd_pipeline(T, B)) else None
        rows0 = (self.WAVEFRONT_CHUNK if hook is not None else T) * B
        ops.gemm(x.view(T * B, I)[:rows0], self.kernels[0].data[:I], ws[0]['gates'].view(T * B, 4 * r0)[:rows0],
                 bias=self.biases[0].data, a_exact=self._binary_inputs, b_weight=True)
        if self._use_wavefront(T, B):
            self._forward_wavefront(ws, outs, T, B, keep, u, seed, dropout, hook, x)
        else:
            tail = chunk_hook is not None and self.use_any_pipeline(T, B)
            for l, r in enumerate(self._num_units):
                w = ws[l]
                kern = self.kernels[l].data
                i_l = self.in_dims()[l]
                if l > 0:
                    ops.gemm(outs[l - 1].view(T * B, i_l), kern[:i_l], w['gates'].view(T * B, 4 * r), bias=self.biases[l].data,
                             b_weight=True)
                if tail and l == self.num_layers - 1:
                    self._forward_top_chunked(w, kern[i_l:], outs, T, B, keep, u, seed, dropout, chunk_hook)
                    continue
                ops.lstm_seq_fwd(w['gates'], kern[i_l:], w['hbuf'], w['cbuf'],
                                 out=w['out'] if dropout else None, dscale=w['dscale'] if dropout else None,
                                 u=None if u is None else u[l], keep=keep, seed=seed + 7919 * l)
        self._saved = (x, ws, dropout)
        state = [LSTMStateTuple(w['cbuf'][T], w['hbuf'][T]) for w in ws]
        return outs[-1], state

    def _forward_top_chunked(self, w, wh, outs, T, B, keep, u, seed, dropout, hook):
        """Large batches (no layer wavefront: the recurrence kernels of two layers do not fit side by side): the TOP
        layer's recurrence leaves SMs free (64 of 148 CTAs at B=2048, R=256), so it runs in time chunks on its own
        stream and the consumer work of the first PIPE_TAIL_STEPS steps runs beside it on the bulk stream; the rest
        follows as one full-width call."""
        C = self.WAVEFRONT_CHUNK
        nch = T // C
        l = self.num_layers - 1
        r = self._num_units[l]
        main = torch.cuda.current_stream()
        st = self._streams(w['gates'].device)[l]
        start = self._event('fwd top start')
        start.record(main)
        n = ops.lstm_seq_ctas(C, B, r, 48)
        bulk_sms = max(16, ops.num_sms() - n)
        t_slow = min(self.PIPE_TAIL_STEPS, C)
        done = None
        for c in range(nch):
            t0, t1 = c * C, (c + 1) * C
            with torch.cuda.stream(st):
                if c == 0:
                    st.wait_event(start)
                ops.set_sm_budget(48)
                try:
                    ops.lstm_seq_fwd(w['gates'][t0:t1], wh, w['hbuf'][t0:t1 + 1], w['cbuf'][t0:t1 + 1],
                                     out=w['out'][t0:t1] if dropout else None,
                                     dscale=w['dscale'][t0:t1] if dropout else None,
                                     u=None if u is None else u[l][t0:t1], keep=keep, seed=seed + 7919 * l, t_base=t0)
                finally:
                    ops.set_sm_budget(0)
                done = self._event(f'fwd top chunk {c}')
                done.record(st)
            if c == 0:
                hook(0, t_slow, done, outs[-1], bulk_sms, False)
        hook(t_slow, T, done, outs[-1], 0, True)
        main.wait_event(done)

    def _forward_wavefront(self, ws, outs, T, B, keep, u, seed, dropout, hook=None, x=None):
        C = self.WAVEFRONT_CHUNK
        nch = T // C
        budgets = self.PIPE_FWD_BUDGETS if hook is not None else (96, 48)
        main = torch.cuda.current_stream()
        streams = self._streams(ws[0]['gates'].device)
        start = self._event('fwd start')
        start.record(main)
        done = [[self._event(f'fwd L{l} chunk {c}') for c in range(nch)] for l in range(self.num_layers)]
        proj = [None] * nch
        n_slow = 0
        if hook is not None:
            # layer-0 input projection of the later chunks: bulk stream, inside the bulk SM budget
            bulk = self.bulk_stream(ws[0]['gates'].device)
            bulk_sms = self.bulk_budget(T, B)
            n_slow = self.PIPE_SLOW_HOOKS if self.PIPE_SLOW_HOOKS > 0 else (2 if B < 512 else (3 if B < 1024 else 1))   # measured, T=256
            n_slow = max(0, min(n_slow * 32 // C, nch - 1))
            I, r0 = self._num_inputs, self._num_units[0]
            with torch.cuda.stream(bulk):
                bulk.wait_event(start)
                ops.set_sm_budget(bulk_sms)
                try:
                    for c in range(1, nch):
                        ops.gemm(x[c * C:(c + 1) * C].view(C * B, I), self.kernels[0].data[:I],
                                 ws[0]['gates'][c * C:(c + 1) * C].view(C * B, 4 * r0), bias=self.biases[0].data,
                                 a_exact=self._binary_inputs, b_weight=True)
                        proj[c] = self._event(f'fwd L0 projection chunk {c}')
                        proj[c].record(bulk)
                finally:
                    ops.set_sm_budget(0)
        for c in range(nch):
            t0, t1 = c * C, (c + 1) * C
            for l, r in enumerate(self._num_units):
                w = ws[l]
                kern = self.kernels[l].data
                i_l = self.in_dims()[l]
                with torch.cuda.stream(streams[l]):
                    if c == 0:
                        streams[l].wait_event(start)
                    if l == 0 and proj[c] is not None:
                        streams[l].wait_event(proj[c])
                    if l > 0:
                        streams[l].wait_event(done[l - 1][c])
                        ops.set_sm_budget(self.WAVEFRONT_GEMM_SMS)
                        try:
                            ops.gemm(outs[l - 1][t0:t1].reshape(C * B, i_l), kern[:i_l],
                                     w['gates'][t0:t1].view(C * B, 4 * r), bias=self.biases[l].data, b_weight=True)
                        finally:
                            ops.set_sm_budget(0)
                    ops.set_sm_budget(budgets[min(l, len(budgets) - 1)])   # wavefront mode: co-resident configurations
                    try:
                        ops.lstm_seq_fwd(w['gates'][t0:t1], kern[i_l:], w['hbuf'][t0:t1 + 1], w['cbuf'][t0:t1 + 1],
                                         out=w['out'][t0:t1] if dropout else None,
                                         dscale=w['dscale'][t0:t1] if dropout else None,
                                         u=None if u is None else u[l][t0:t1], keep=keep,
                                         seed=seed + 7919 * l, t_base=t0)
                    finally:
                        ops.set_sm_budget(0)
                    done[l][c].record(streams[l])
            if hook is not None and c < n_slow:
                hook(t0, t1, done[self.num_layers - 1][c], outs[-1], bulk_sms, False)
        if hook is not None:
            top_done = done[self.num_layers - 1][nch - 1]
            if self.PIPE_DESCENDING and self.use_pipeline(T, B) and nch - n_slow > 1:
                beside = self.bulk_budget(T, B, backward=True)
                for c in reversed(range(n_slow, nch)):
                    hook(c * C, (c + 1) * C, top_done, outs[-1], 0 if c == nch - 1 else beside, c == n_slow)
            else:
                hook(n_slow * C, T, top_done, outs[-1], 0, True)
        for l in range(self.num_layers):
            main.wait_event(done[l][nch - 1])

    def backward_sequence(self, dout, need_dx=False, pipelined=False, dout_ready=None):
        """dout[T,B,R_top] = grad wrt the (dropped-out) top outputs. Writes kernel/bias grads; returns dx or None.
        pipelined: weight / bias gradients are ACCUMULATED chunk by chunk on the bulk stream while BPTT walks the
        earlier chunks (the caller zeroed the gradient buffers, as GradientApplier.zero_grad does every step)."""
        x, ws, dropout = self._saved
        T, B, I = x.shape
        for l, r in enumerate(self._num_units):
            w = ws[l]
            w.setdefault('dh_work', torch.empty(B, r, device=x.device))
            w.setdefault('dc_work', torch.empty(B, r, device=x.device))
            i_l = self.in_dims()[l]
            if (l > 0 or need_dx) and (w.get('d_in') is None or w['d_in'].shape != (T, B, i_l)):
                w['d_in'] = torch.empty(T, B, i_l, device=x.device)
        pipelined = pipelined and self.use_pipeline(T, B)
        if self._use_wavefront(T, B, backward=True):
            self._backward_wavefront(ws, dout, T, B, dropout, need_dx, x if pipelined else None,
                                     dout_ready if pipelined else None)
            if pipelined:
                return ws[0]['d_in'] if need_dx else None
        else:
            d = dout
            for l in reversed(range(self.num_layers)):
                w = ws[l]
                kern = self.kernels[l]
                i_l = self.in_dims()[l]
                ops.lstm_seq_bwd(w['gates'], kern.data[i_l:], w['cbuf'], d, w['dscale'] if dropout else None,
                                 w['dh_work'], w['dc_work'])
                if l > 0 or need_dx:
                    ops.gemm(w['gates'].view(T * B, -1), kern.data[:i_l], w['d_in'].view(T * B, i_l), transB=True,
                             b_weight=True)   # dx = dG Wx^T
                    d = w['d_in']
        # weight gradients: batched GEMMs over all T*B rows
        side = self.aux_stream(x.device) if (self.COLSUM_SIDE_STREAM and x.is_cuda) else None
        for l in reversed(range(self.num_layers)):
            self._weight_grads(l, x, ws, dropout, 0, T, B, beta=0.0, side=side)
        if side is not None:
            ev = torch.cuda.Event()
            ev.record(side)
            torch.cuda.current_stream().wait_event(ev)
        return ws[0]['d_in'] if need_dx else None

    def _weight_grads(self, l, x, ws, dropout, t0, t1, B, beta, side=None):
        """dWx_l (+)= in_l^T dG_l, dWh_l (+)= h_{l,t-1}^T dG_l, db_l (+)= colsum(dG_l) over steps [t0, t1)."""
        r = self._num_units[l]
        w = ws[l]
        kern = self.kernels[l]
        i_l = self.in_dims()[l]
        rows = (t1 - t0) * B
        dg = w['gates'][t0:t1].view(rows, 4 * r)                       # now d(pre-activations)
        inp = x[t0:t1].view(rows, -1) if l == 0 else \
            (ws[l - 1]['out'] if dropout else ws[l - 1]['hbuf'][1:])[t0:t1].reshape(rows, i_l)
        if side is not None:
            # the bias-gradient column sums (HBM-bound, no shared memory) run UNDER the weight-gradient GEMMs (tensor /
            # shared-memory bound) on a side stream; both read the same dG, so one of them mostly hits L2
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                side.wait_event(ev)
                ops.colsum(dg, self.biases[l].grad, accumulate=beta != 0.0)
        ops.gemm(inp, dg, kern.grad[:i_l], transA=True, beta=beta, a_exact=(l == 0 and self._binary_inputs))
        ops.gemm(w['hbuf'][t0:t1].view(rows, r), dg, kern.grad[i_l:], transA=True, beta=beta)
        if side is None:
            ops.colsum(dg, self.biases[l].grad, accumulate=beta != 0.0)

    def _backward_wavefront(self, ws, dout, T, B, dropout, need_dx, x_pipe=None, dout_ready=None):
        """BPTT as a wavefront over time chunks: layer l back-propagates chunk c while layer l-1 works on chunk c+1.
        x_pipe (pipeline mode): the layer-0 inputs; weight gradients of finished chunks run on the bulk stream."""
        C = self.WAVEFRONT_CHUNK
        nch = T // C
        L = self.num_layers
        main = torch.cuda.current_stream()
        streams = self._streams(ws[0]['gates'].device)
        start = self._event('bwd start')
        start.record(main)
        done = [[self._event(f'bwd L{l} chunk {c}') for c in range(nch)] for l in range(L)]
        budgets = [0] * L
        if L == 2:
            budgets = [72, 36]          # SMs for the persistent BPTT kernels of the two layers (they must co-reside)
            if x_pipe is not None:
                budgets = list(self._bwd_budgets(B))
        bulk = self.bulk_stream(ws[0]['gates'].device) if x_pipe is not None else None
        for c in reversed(range(nch)):
            t0, t1 = c * C, (c + 1) * C
            for l in reversed(range(L)):
                w = ws[l]
                kern = self.kernels[l]
                i_l = self.in_dims()[l]
                r = self._num_units[l]
                d = dout if l == L - 1 else ws[l + 1]['d_in']
                aux = self.aux_stream(w['gates'].device) if (bulk is not None and self.PIPE_AUX_SMS > 0) else None
                cell_done = done[l][c]
                with torch.cuda.stream(streams[l]):
                    if c == nch - 1:
                        streams[l].wait_event(start)
                    if l < L - 1:
                        streams[l].wait_event(done[l + 1][c])
                    elif dout_ready is not None:
                        for ev in dout_ready.get(c, ()):        # the chunk's grad wrt the top outputs (bulk stream)
                            streams[l].wait_event(ev)
                    ops.set_sm_budget(budgets[l])
                    try:
                        ops.lstm_seq_bwd(w['gates'][t0:t1], kern.data[i_l:], w['cbuf'][t0:t1 + 1], d[t0:t1],
                                         w['dscale'][t0:t1] if dropout else None, w['dh_work'], w['dc_work'],
                                         has_next=c < nch - 1)
                        if (l > 0 or need_dx) and aux is None:
                            ops.set_sm_budget(self.WAVEFRONT_GEMM_SMS)
                            ops.gemm(w['gates'][t0:t1].view(C * B, 4 * r), kern.data[:i_l],
                                     w['d_in'][t0:t1].view(C * B, i_l), transB=True, b_weight=True)
                    finally:
                        ops.set_sm_budget(0)
                    if (l > 0 or need_dx) and aux is not None:
                        cell_done = self._event(f'bwd L{l} cells chunk {c}')
                        cell_done.record(streams[l])
                    else:
                        done[l][c].record(streams[l])
                if (l > 0 or need_dx) and aux is not None:
                    # dx = dG Wx^T of the chunk on its own stream: the layer's next BPTT chunk does not wait for it
                    with torch.cuda.stream(aux):
                        aux.wait_event(cell_done)
                        ops.set_sm_budget(self.PIPE_AUX_SMS)
                        try:
                            ops.gemm(w['gates'][t0:t1].view(C * B, 4 * r), kern.data[:i_l],
                                     w['d_in'][t0:t1].view(C * B, i_l), transB=True, b_weight=True)
                        finally:
                            ops.set_sm_budget(0)
                        done[l][c].record(aux)
                if bulk is not None:
                    with torch.cuda.stream(bulk):
                        bulk.wait_event(cell_done)
                        ops.set_sm_budget(0 if c < max(1, self.PIPE_FULL_WGRADS * 32 // C) else self.bulk_budget(T, B, backward=True))
                        try:
                            self._weight_grads(l, x_pipe, ws, dropout, t0, t1, B, beta=1.0)
                        finally:
                            ops.set_sm_budget(0)
                        if self.TRACE is not None:
                            self._event(f'bwd wgrads L{l} chunk {c}').record(bulk)
        for l in range(L):
            main.wait_event(done[l][0])
        if bulk is not None:
            ev = self._event('bwd bulk end')
            ev.record(bulk)
            main.wait_event(ev)

    # ------------------------------------------------------------------ one step (generation)
    def step(self, x, state, scratch=None):
        """One MultiRNNCell step without dropout (is_train=False): x[B,I], state -> (out[B,R_top], new state).
        rnn.py:178-194 called from generators/rnn_nade.py:267."""
        B = x.shape[0]
        new_state = []
        inp = x
        for l, r in enumerate(self._num_units):
            kern, bias = self.kernels[l].data, self.biases[l].data
            i_l = inp.shape[1]
            g = torch.empty(B, 4 * r, device=x.device)
            ops.gemm(inp, kern[:i_l], g, bias=bias)
            ops.gemm(state[l][1], kern[i_l:], g, beta=1.0)
            c = torch.empty(B, r, device=x.device)
            h = torch.empty(B, r, device=x.device)
            ops.lstm_cell_fwd(g, state[l][0], c, h)
            new_state.append(LSTMStateTuple(c, h))
            inp = h
        return inp, new_state
