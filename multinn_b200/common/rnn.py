"""Temporal unit: stacked LSTM with output dropout (mirrors reference models/common/rnn.py:13-219).

Cell semantics = CudnnCompatibleLSTMCell == LSTMBlockCell(forget_bias=0): kernel[(I+R),4R] with rows
[x ; h] and gate column blocks i, j, f, o; DropoutWrapper(output_keep_prob) on every layer's output
(state h is not dropped); MultiRNNCell stacking (rnn.py:104-145).

B200 layout: sequences are time-major [T,B,*] so that one step's rows are contiguous; the input projection
x.Wx + b of a whole sequence is ONE GEMM over all T*B rows (hoisted out of the recurrence), only h.Wh is
sequential. Gate pre-activations are overwritten by their activations (saved for BPTT) and then by their
gradients, so the recurrence keeps a single [T,B,4R] buffer per layer.
"""
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
    def forward_sequence(self, x, keep=1.0, u=None, seed=0, initial_state=None):
        """x[T,B,I] time-major -> outputs[T,B,R_top] (dropped out when keep < 1), final state.
        u: optional list (per layer) of [T,B,R_l] uniforms for reproducible dropout; else Philox(seed).
        Saves what BPTT needs (call `backward_sequence` next)."""
        T, B, I = x.shape
        assert I == self._num_inputs and x.is_contiguous()
        dropout = keep < 1.0
        ws = self._workspace(T, B, x.device, dropout)
        inp = x.view(T * B, I)
        for l, r in enumerate(self._num_units):
            w = ws[l]
            kern, bias = self.kernels[l].data, self.biases[l].data
            i_l = inp.shape[1]
            gates = w['gates'].view(T * B, 4 * r)
            ops.gemm(inp, kern[:i_l], gates, bias=bias, a_exact=(l == 0 and self._binary_inputs))   # hoisted input projection
            if initial_state is None:
                w['hbuf'][0].zero_()
                w['cbuf'][0].zero_()
            else:
                w['cbuf'][0].copy_(initial_state[l][0])
                w['hbuf'][0].copy_(initial_state[l][1])
            ops.lstm_seq_fwd(w['gates'], kern[i_l:], w['hbuf'], w['cbuf'],
                             out=w['out'] if dropout else None, dscale=w['dscale'] if dropout else None,
                             u=None if u is None else u[l], keep=keep, seed=seed + 7919 * l)
            out = w['out'] if dropout else w['hbuf'][1:]
            inp = out.view(T * B, r)
        self._saved = (x, ws, dropout)
        state = [LSTMStateTuple(w['cbuf'][T], w['hbuf'][T]) for w in ws]
        return out, state

    def backward_sequence(self, dout, need_dx=False):
        """dout[T,B,R_top] = grad wrt the (dropped-out) top outputs. Writes kernel/bias grads; returns dx or None."""
        x, ws, dropout = self._saved
        T, B, I = x.shape
        dx = None
        for l in reversed(range(self.num_layers)):
            r = self._num_units[l]
            w = ws[l]
            kern = self.kernels[l]
            i_l = self.in_dims()[l]
            dh_work = w.setdefault('dh_work', torch.empty(B, r, device=x.device))
            dc_work = w.setdefault('dc_work', torch.empty(B, r, device=x.device))
            ops.lstm_seq_bwd(w['gates'], kern.data[i_l:], w['cbuf'], dout, w['dscale'] if dropout else None,
                             dh_work, dc_work)
            dg = w['gates'].view(T * B, 4 * r)                       # now d(pre-activations)
            inp = x.view(T * B, I) if l == 0 else \
                (ws[l - 1]['out'] if dropout else ws[l - 1]['hbuf'][1:]).view(T * B, i_l)
            ops.gemm(inp, dg, kern.grad[:i_l], transA=True, a_exact=(l == 0 and self._binary_inputs))   # dWx = x^T dG
            ops.gemm(w['hbuf'][:T].view(T * B, r), dg, kern.grad[i_l:], transA=True)   # dWh = h_{t-1}^T dG
            ops.colsum(dg, self.biases[l].grad)
            if l > 0 or need_dx:
                d_in = w.get('d_in')
                if d_in is None or d_in.shape != (T, B, i_l):
                    d_in = w['d_in'] = torch.empty(T, B, i_l, device=x.device)
                ops.gemm(dg, kern.data[:i_l], d_in.view(T * B, i_l), transB=True)    # dx = dG Wx^T
                dout = d_in
                if l == 0:
                    dx = d_in
        return dx

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
