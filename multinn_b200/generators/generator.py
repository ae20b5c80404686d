"""Abstract Generator (mirrors reference models/generators/generator.py:9-205)."""
import abc

from ..common.model import Model


class Generator(Model, abc.ABC):
    def __init__(self, num_dims, num_hidden, num_hidden_rnn, keep_prob=1.0, internal_bias=False, name='generator',
                 track_name='all'):
        super().__init__(name=name)
        self._track_name = track_name
        self._num_dims = num_dims
        self._num_hidden = [num_hidden] if isinstance(num_hidden, int) else list(num_hidden)
        self._num_hidden_rnn = [num_hidden_rnn] if isinstance(num_hidden_rnn, int) else list(num_hidden_rnn)
        self._keep_prob = keep_prob
        self._internal_bias = internal_bias
        self._lengths = None

    num_dims = property(lambda s: s._num_dims)
    num_hidden = property(lambda s: s._num_hidden)
    num_hidden_rnn = property(lambda s: s._num_hidden_rnn)
    track_name = property(lambda s: s._track_name)
    keep_prob = property(lambda s: s._keep_prob)
    internal_bias = property(lambda s: s._internal_bias)

    @abc.abstractmethod
    def zero_state(self, batch_size):
        ...

    @abc.abstractmethod
    def generate(self, x, num_steps):
        ...

    def pretrain(self, inputs_flat, lr, u=None, seed=None):
        """generator.py:125-150: pre-training of a generator module on the flattened inputs [rows, num_dims]. The default
        is the RNN-NADE behaviour (rnn_nade.py:320-326: nothing to pre-train, no update) and returns None."""
        return None

