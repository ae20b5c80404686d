"""RNN-Estimator base (mirrors reference models/generators/rnn_estimator.py:12-323): a temporal RNN whose
per-step outputs parameterise a distribution estimator's biases."""
import abc
import collections

import torch

from ..common.rnn import RNN
from .generator import Generator

_Base = collections.namedtuple('RnnEstimatorStateTuple', ('b_enc', 'b_dec', 'rnn_state'))


class RnnEstimatorStateTuple(_Base):
    """(b_enc, b_dec, rnn_state) -- rnn_estimator.py:12-36. For RnnMultiNADE b_enc/b_dec are lists of M tensors;
    all of them are column views of one Dense output buffer `fc` (kept in `.fc` when produced here)."""
    __slots__ = ()


class RnnEstimator(Generator, abc.ABC):
    def __init__(self, arena, num_inputs, num_dims, num_hidden, num_hidden_rnn, keep_prob=1.0, internal_bias=False,
                 name='rnn-estimator', track_name='all', binary_inputs=False):
        super().__init__(num_dims, num_hidden, num_hidden_rnn, keep_prob, internal_bias, name, track_name)
        self._arena = arena
        self._num_inputs = num_inputs
        self._rnn = RNN(arena, num_inputs, self._num_hidden_rnn, keep_prob=keep_prob, name=f'{name}/rnn',
                        binary_inputs=binary_inputs)
        self._init_estimator()

    @abc.abstractmethod
    def _init_estimator(self):
        ...

    @property
    def num_inputs(self):
        return self._num_inputs

    @property
    def rnn(self):
        return self._rnn

    def _get_rnn_zero_state(self, batch_size, device):
        return self._rnn.zero_state(batch_size, device)

    @abc.abstractmethod
    def _get_state(self, inputs, lengths=None, initial_state=None, last_outputs=False):
        ...

    def steps(self, inputs, initial_state=None):
        """rnn_estimator.py:222-235."""
        return self._get_state(inputs, initial_state=initial_state, last_outputs=True)

    @abc.abstractmethod
    def single_step(self, inputs, initial_state):
        ...

    @abc.abstractmethod
    def sample_single(self, inputs, state, u=None):
        ...

    def generate(self, x, num_steps, u=None, seed=0):
        """rnn_estimator.py:271-323. x[T,B,I] TIME-MAJOR intro (incl. the leading zero frame) ->
        samples[B,num_steps,num_outputs]. u[num_steps, M, B, D] optional uniforms (else Philox)."""
        state = self._get_state(x, lengths=None, last_outputs=True)
        intro = x[-1]
        B = x.shape[1]
        out = torch.empty(B, num_steps, self.num_outputs, device=x.device)
        for s in range(num_steps):
            samples, _ = self.sample_single(intro, state, u=None if u is None else u[s], seed=seed, offset=s,
                                            out=out[:, s])
            state = self.single_step(samples, state)
            intro = samples
        return out
