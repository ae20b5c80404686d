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
    WAVEFRONT_CHUNK = 32
    WAVEFRONT_GEMM_SMS = 32

    def _use_wavefront(self, T, B, backward=False):
        mb = min(self.WAVEFRONT_MAX_BATCH, self.WAVEFRONT_MAX_BATCH_BWD) if backward else self.WAVEFRONT_MAX_BATCH
        return (self.num_layers > 1 and B <= mb and T % self.WAVEFRONT_CHUNK == 0 and T >= 2 * self.WAVEFRONT_CHUNK)

    def _streams(self, device):
        st = self.__dict__.get('_wave_streams')
        if st is None:
            st = self._wave_streams = [torch.cuda.Stream(device=device) for _ in range(self.num_layers)]
        return st

    def forward_sequence(self, x, keep=1.0, u=None, seed=0, initial_state=None):
        """x[T,B,I] time-major -> outputs[T,B,R_top] (dropped out when keep < 1), final state.
        u: optional list (per layer) of [T,B,R_l] uniforms for reproducible dropout; else Philox(seed).
        Saves what BPTT needs (call `backward_sequence` next)."""
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
        # layer 0: hoisted input projection over all T*B rows
        r0 = self._num_units[0]
        ops.gemm(x.view(T * B, I), self.kernels[0].data[:I], ws[0]['gates'].view(T * B, 4 * r0), bias=self.biases[0].data,
                 a_exact=self._binary_inputs)
        if self._use_wavefront(T, B):
            self._forward_wavefront(ws, outs, T, B, keep, u, seed, dropout)
        else:
            for l, r in enumerate(self._num_units):
                w = ws[l]
                kern = self.kernels[l].data
                i_l = self.in_dims()[l]
                if l > 0:
                    ops.gemm(outs[l - 1].view(T * B, i_l), kern[:i_l], w['gates'].view(T * B, 4 * r), bias=self.biases[l].data)
                ops.lstm_seq_fwd(w['gates'], kern[i_l:], w['hbuf'], w['cbuf'],
                                 out=w['out'] if dropout else None, dscale=w['dscale'] if dropout else None,
                                 u=None if u is None else u[l], keep=keep, seed=seed + 7919 * l)
        self._saved = (x, ws, dropout)
        state = [LSTMStateTuple(w['cbuf'][T], w['hbuf'][T]) for w in ws]
        return outs[-1], state

    def _forward_wavefront(self, ws, outs, T, B, keep, u, seed, dropout):
        C = self.WAVEFRONT_CHUNK
        nch = T // C
        main = torch.cuda.current_stream()
        streams = self._streams(ws[0]['gates'].device)
        start = torch.cuda.Event()
        start.record(main)
        done = [[torch.cuda.Event() for _ in range(nch)] for _ in range(self.num_layers)]
        for c in range(nch):
            t0, t1 = c * C, (c + 1) * C
            for l, r in enumerate(self._num_units):
                w = ws[l]
                kern = self.kernels[l].data
                i_l = self.in_dims()[l]
                with torch.cuda.stream(streams[l]):
                    if c == 0:
                        streams[l].wait_event(start)
                    if l > 0:
                        streams[l].wait_event(done[l - 1][c])
                        ops.set_sm_budget(self.WAVEFRONT_GEMM_SMS)
                        try:
                            ops.gemm(outs[l - 1][t0:t1].reshape(C * B, i_l), kern[:i_l],
                                     w['gates'][t0:t1].view(C * B, 4 * r), bias=self.biases[l].data)
                        finally:
                            ops.set_sm_budget(0)
                    ops.set_sm_budget(96 if l == 0 else 48)     # wavefront mode: kernels pick co-resident configurations
                    try:
                        ops.lstm_seq_fwd(w['gates'][t0:t1], kern[i_l:], w['hbuf'][t0:t1 + 1], w['cbuf'][t0:t1 + 1],
                                         out=w['out'][t0:t1] if dropout else None,
                                         dscale=w['dscale'][t0:t1] if dropout else None,
                                         u=None if u is None else u[l][t0:t1], keep=keep,
                                         seed=seed + 7919 * l + 104729 * c)
                    finally:
                        ops.set_sm_budget(0)
                    done[l][c].record(streams[l])
        for l in range(self.num_layers):
            main.wait_event(done[l][nch - 1])

    def backward_sequence(self, dout, need_dx=False):
        """dout[T,B,R_top] = grad wrt the (dropped-out) top outputs. Writes kernel/bias grads; returns dx or None."""
        x, ws, dropout = self._saved
        T, B, I = x.shape
        for l, r in enumerate(self._num_units):
            w = ws[l]
            w.setdefault('dh_work', torch.empty(B, r, device=x.device))
            w.setdefault('dc_work', torch.empty(B, r, device=x.device))
            i_l = self.in_dims()[l]
            if (l > 0 or need_dx) and (w.get('d_in') is None or w['d_in'].shape != (T, B, i_l)):
                w['d_in'] = torch.empty(T, B, i_l, device=x.device)
        if self._use_wavefront(T, B, backward=True):
            self._backward_wavefront(ws, dout, T, B, dropout, need_dx)
        else:
            d = dout
            for l in reversed(range(self.num_layers)):
                w = ws[l]
                kern = self.kernels[l]
                i_l = self.in_dims()[l]
                ops.lstm_seq_bwd(w['gates'], kern.data[i_l:], w['cbuf'], d, w['dscale'] if dropout else None,
                                 w['dh_work'], w['dc_work'])
                if l > 0 or need_dx:
                    ops.gemm(w['gates'].view(T * B, -1), kern.data[:i_l], w['d_in'].view(T * B, i_l), transB=True)   # dx = dG Wx^T
                    d = w['d_in']
        # weight gradients: batched GEMMs over all T*B rows
        for l in reversed(range(self.num_layers)):
            r = self._num_units[l]
            w = ws[l]
            kern = self.kernels[l]
            i_l = self.in_dims()[l]
            dg = w['gates'].view(T * B, 4 * r)                       # now d(pre-activations)
            inp = x.view(T * B, I) if l == 0 else \
                (ws[l - 1]['out'] if dropout else ws[l - 1]['hbuf'][1:]).view(T * B, i_l)
            ops.gemm(inp, dg, kern.grad[:i_l], transA=True, a_exact=(l == 0 and self._binary_inputs))   # dWx = x^T dG
            ops.gemm(w['hbuf'][:T].view(T * B, r), dg, kern.grad[i_l:], transA=True)   # dWh = h_{t-1}^T dG
            ops.colsum(dg, self.biases[l].grad)
        return ws[0]['d_in'] if need_dx else None

    def _backward_wavefront(self, ws, dout, T, B, dropout, need_dx):
        """BPTT as a wavefront over time chunks: layer l back-propagates chunk c while layer l-1 works on chunk c+1."""
        C = self.WAVEFRONT_CHUNK
        nch = T // C
        L = self.num_layers
        main = torch.cuda.current_stream()
        streams = self._streams(ws[0]['gates'].device)
        start = torch.cuda.Event()
        start.record(main)
        done = [[torch.cuda.Event() for _ in range(nch)] for _ in range(L)]
        budgets = [0] * L
        if L == 2:
            budgets = [72, 36]          # SMs for the persistent BPTT kernels of the two layers (they must co-reside)
        for c in reversed(range(nch)):
            t0, t1 = c * C, (c + 1) * C
            for l in reversed(range(L)):
                w = ws[l]
                kern = self.kernels[l]
                i_l = self.in_dims()[l]
                r = self._num_units[l]
                d = dout if l == L - 1 else ws[l + 1]['d_in']
                with torch.cuda.stream(streams[l]):
                    if c == nch - 1:
                        streams[l].wait_event(start)
                    if l < L - 1:
                        streams[l].wait_event(done[l + 1][c])
                    ops.set_sm_budget(budgets[l])
                    try:
                        ops.lstm_seq_bwd(w['gates'][t0:t1], kern.data[i_l:], w['cbuf'][t0:t1 + 1], d[t0:t1],
                                         w['dscale'][t0:t1] if dropout else None, w['dh_work'], w['dc_work'],
                                         has_next=c < nch - 1)
                        if l > 0 or need_dx:
                            ops.set_sm_budget(self.WAVEFRONT_GEMM_SMS)
                            ops.gemm(w['gates'][t0:t1].view(C * B, 4 * r), kern.data[:i_l],
                                     w['d_in'][t0:t1].view(C * B, i_l), transB=True)
                    finally:
                        ops.set_sm_budget(0)
                    done[l][c].record(streams[l])
        for l in range(L):
            main.wait_event(done[l][0])

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
