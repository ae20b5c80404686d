"""NADE distribution estimator (mirrors reference models/common/nade.py:20-329).

`NADE.log_prob(x, b_enc, b_dec) -> (nll[N], cond_p[N,D])` and `NADE.sample(b_enc, b_dec, n, temperature)
-> (v[N,D], nll[N])` keep the reference signatures (nade.py:155-229, :231-308). The arithmetic runs in
the sm_100a kernels of csrc/nade.cu; M NADEs of one MultiNADE share a weight bank [M,D,H] so that all
tracks go through one launch (rnn_multinade.py:58-71 keeps a python list of M NADE objects instead).
"""
import math

import torch

from .. import ops
from ..params import truncated_normal


class NADEBank:
    """Weights of M NADEs: w_enc[M,D,H] (reference per-NADE w_enc[D,1,H]) and w_dec[M,D,H] (w_dec[D,H,1])."""

    def __init__(self, arena, num_tracks, num_dims, num_hidden, name='nade'):
        if num_dims > 128:
            # fail at construction, not at the first launch: the kernels keep both weight matrices of a track in shared
            # memory and the row masks in four words. `training.num_pixels: 3` of the reference's shipped
            # default config gives num_dims = 252: use `num_pixels: 1` (the BASELINE shapes) with this package
            raise NotImplementedError(f'NADE kernels are built for num_dims <= 128 (and num_hidden 128 or 256); got '
                                      f'num_dims={num_dims}')
        std = 1.0 / math.sqrt(num_dims)                       # nade.py:48-50
        self.num_tracks, self.num_dims, self.num_hidden = num_tracks, num_dims, num_hidden
        self.w_enc = arena.add(f'{name}/w_enc', (num_tracks, num_dims, num_hidden), truncated_normal(std))
        self.w_dec = arena.add(f'{name}/w_dec', (num_tracks, num_dims, num_hidden), truncated_normal(std))

    @property
    def trainable_params(self):
        return [self.w_enc, self.w_dec]


class NADE:
    """One NADE (a view of track `track` of a bank, or its own bank of one)."""

    def __init__(self, num_dims, num_hidden=128, internal_bias=False, name='nade', arena=None, bank=None, track=0):
        if internal_bias:
            # All reference call sites (generators/rnn_nade.py:59-62, rnn_multinade.py:66-70 with the modes'
            # internal_bias=False) feed external biases.
            raise NotImplementedError('NADE(internal_bias=True) is outside the B200 hot path')
        self.name = name
        self._num_dims, self._num_hidden, self._internal_bias = num_dims, num_hidden, internal_bias
        if bank is None:
            assert arena is not None, 'NADE needs a ParamArena (or a NADEBank)'
            bank = NADEBank(arena, 1, num_dims, num_hidden, name=name)
        self._bank, self._track = bank, track

    num_dims = property(lambda s: s._num_dims)
    num_hidden = property(lambda s: s._num_hidden)
    internal_bias = property(lambda s: s._internal_bias)

    @property
    def w_enc(self):
        return self._bank.w_enc.data[self._track]

    @property
    def w_dec(self):
        return self._bank.w_dec.data[self._track]

    def _fc(self, b_enc, b_dec, n=None):
        if b_enc is None or b_dec is None:
            raise ValueError('Bias values should be provided when `internal_bias` is `False`')   # nade.py:180-181
        N = n or max(b_enc.shape[0], b_dec.shape[0])
        H, D = self._num_hidden, self._num_dims
        ld = (H + D + 3) // 4 * 4
        fc = torch.empty(N, ld, device=b_enc.device)
        fc[:, :H] = b_enc            # broadcasts a [1,H] bias like the tf.tile at nade.py:184-187
        fc[:, H:H + D] = b_dec
        return fc

    def log_prob(self, x, b_enc=None, b_dec=None):
        """x[N,D] in {0,1} -> (nll[N] positive, cond_p[N,D]) (nade.py:155-229)."""
        fc = self._fc(b_enc, b_dec, x.shape[0])
        N, D, H = x.shape[0], self._num_dims, self._num_hidden
        bits = torch.empty(N, 4, dtype=torch.int32, device=x.device)
        ops.pack_rows(x.contiguous(), bits, D)
        nll = torch.empty(1, N, device=x.device)
        cond_p = torch.empty(1, N, D, device=x.device)
        ops.nade_logprob_fwd(bits, fc, 0, H, self.w_enc.unsqueeze(0), self.w_dec.unsqueeze(0), nll, cond_p)
        return nll[0], cond_p[0]

    def sample(self, b_enc=None, b_dec=None, n=None, temperature=None, u=None, seed=0, offset=0):
        """(v[N,D], nll[N]) (nade.py:231-308). temperature=None -> threshold p >= .5; temperature=1. -> Bernoulli
        sampling v_i = float(u_i < p_i) with supplied uniforms `u[N,D]` or in-kernel Philox(seed, offset)."""
        if temperature is not None and float(temperature) != 1.0:
            raise NotImplementedError('only temperature in (None, 1.0) is used by the reference call sites')
        fc = self._fc(b_enc, b_dec, n)
        N, D, H = fc.shape[0], self._num_dims, self._num_hidden
        v = torch.empty(N, D, device=fc.device)
        nll = torch.empty(1, N, device=fc.device)
        sampling = temperature is not None
        ops.nade_sample(fc, 0, H, self.w_enc.unsqueeze(0), self.w_dec.unsqueeze(0), v, D, 1, 0,
                        u=u.contiguous().view(1, N, D) if (sampling and u is not None) else None,
                        use_philox=sampling and u is None, seed=seed, offset=offset, nll=nll)
        return v, nll[0]
