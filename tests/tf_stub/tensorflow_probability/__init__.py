"""TEST INFRASTRUCTURE ONLY -- stand-in for tensorflow_probability 0.6.0's Bernoulli (the one TFP name the reference
uses: models/common/nade.py:283-287, models/common/rbm.py:375-387). TFP 0.6.0 `Bernoulli._sample_n` draws
`uniform = random_uniform(shape)` and returns `cast(uniform < probs, dtype)` with probs = sigmoid(logits) when built
from logits: a STRICT less-than. Here the uniforms come from an injected stream so that the reference's sampling code
can be replayed deterministically: `push_uniforms(list_of_arrays)` queues one array per `.sample()` call, in call order
(each is consumed once; its shape must match the distribution's batch shape)."""
import collections
import types

import numpy as np
from scipy.special import expit as _expit

import tensorflow as tf

_queue = collections.deque()
consumed = []          # shapes of the uniform arrays consumed so far (for the generator script's bookkeeping)


def push_uniforms(arrays):
    _queue.extend(np.asarray(a, dtype=np.float64) for a in arrays)


_auto = [None]
auto_log = []          # the uniform arrays drawn by the seeded fallback, in call order (float32-representable)


def auto_uniforms(seed):
    """When the injected queue is empty, draw from a seeded generator instead of raising, and log the arrays (for graphs
    whose draw order is discovered by running them). seed=None switches the fallback off."""
    _auto[0] = None if seed is None else np.random.default_rng(seed)
    del auto_log[:]


def pending():
    return len(_queue)


def clear():
    _queue.clear()
    del consumed[:]


class _Bernoulli:
    def __init__(self, logits=None, probs=None, dtype=None, **kw):
        if (logits is None) == (probs is None):
            raise ValueError('Must pass probs or logits, but not both.')
        self.probs = np.asarray(probs, dtype=np.float64) if probs is not None else _expit(np.asarray(logits, dtype=np.float64))

    def sample(self, sample_shape=(), seed=None, name='sample'):
        if not _queue and _auto[0] is not None:
            u = _auto[0].random(self.probs.shape, dtype=np.float32).astype(np.float64)
            auto_log.append(u)
        elif not _queue:
            raise RuntimeError('tfp stub: Bernoulli.sample() called with no injected uniforms left (push_uniforms)')
        else:
            u = _queue.popleft()
        if u.shape != self.probs.shape:
            raise ValueError(f'tfp stub: injected uniforms have shape {u.shape}, the distribution {self.probs.shape}')
        consumed.append(u.shape)
        return tf._t((u < self.probs).astype(np.float64))


distributions = types.SimpleNamespace(Bernoulli=_Bernoulli)
__version__ = '0.6.0-numpy-stub'
