"""Flat fp32 parameter arena: every trainable tensor of a model is a view into ONE contiguous buffer, with a
matching gradient buffer and Adam slot buffers, so that the data-parallel exchange is one NCCL allreduce and
the optimiser one fused clip+Adam launch over the union of variables (reference utils/training.py:151-177
clips over the whole variable list; train.py:61-64 builds the optimiser).
"""
import math

import numpy as np
import torch


class Param:
    """Handle to one trainable tensor; `.data` / `.grad` are views into the arena once finalised."""

    def __init__(self, name, shape, init):
        self.name, self.shape, self.init = name, tuple(int(s) for s in shape), init
        self.offset = None
        self.data = None
        self.grad = None

    @property
    def numel(self):
        return int(np.prod(self.shape))


def glorot_uniform(fan_in, fan_out):
    """tf.contrib.layers.xavier_initializer / TF default glorot_uniform (SURVEY 8(d) 'Weights')."""
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return lambda rng, shape: rng.uniform(-lim, lim, size=shape).astype(np.float32)


def truncated_normal(std):
    """tf.truncated_normal_initializer(stddev): resample |z| > 2 (reference common/nade.py:49-50)."""

    def f(rng, shape):
        z = rng.standard_normal(shape)
        bad = np.abs(z) > 2
        while bad.any():
            z[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(z) > 2
        return (z * std).astype(np.float32)

    return f


def zeros():
    return lambda rng, shape: np.zeros(shape, np.float32)


class ParamArena:
    """Collects Param specs in creation order, then allocates flat / grad (/ m, v) buffers on a device."""

    ALIGN = 64  # elements; keeps every view 256-byte aligned (vectorised loads, TMA)

    def __init__(self):
        self.params = []
        self.flat = self.grad = self.m = self.v = None
        self.size = 0

    def add(self, name, shape, init):
        p = Param(name, shape, init)
        self.params.append(p)
        return p

    def finalize(self, device, seed=23):
        rng = np.random.default_rng(seed)
        off = 0
        for p in self.params:
            p.offset = off
            off += (p.numel + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.size = off
        host = np.zeros(off, np.float32)
        for p in self.params:
            host[p.offset:p.offset + p.numel] = p.init(rng, p.shape).reshape(-1)
        self.flat = torch.from_numpy(host).to(device)
        self.grad = torch.zeros_like(self.flat)
        for p in self.params:
            p.data = self.flat[p.offset:p.offset + p.numel].view(p.shape)
            p.grad = self.grad[p.offset:p.offset + p.numel].view(p.shape)
        return self

    def ensure_slots(self):
        if self.m is None:
            self.m = torch.zeros_like(self.flat)
            self.v = torch.zeros_like(self.flat)

    def subset(self, params):
        """(offset, length) of the smallest contiguous range covering `params` (they must be adjacent)."""
        lo = min(p.offset for p in params)
        hi = max(p.offset + (p.numel + self.ALIGN - 1) // self.ALIGN * self.ALIGN for p in params)
        return lo, hi - lo

    def named(self):
        return {p.name: p for p in self.params}

    def load(self, name, array):
        """Host -> device copy of one tensor (tests, checkpoint import)."""
        p = self.named()[name]
        t = torch.as_tensor(np.asarray(array, dtype=np.float32)).reshape(p.shape)
        p.data.copy_(t.to(p.data.device))

    def state_dict(self):
        return {p.name: p.data.detach().cpu().clone() for p in self.params}

    def load_state_dict(self, sd):
        for p in self.params:
            if p.name in sd:
                p.data.copy_(sd[p.name].to(p.data.device))
