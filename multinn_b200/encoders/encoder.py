"""Abstract Encoder (mirrors reference models/encoders/encoder.py:8-166)."""
import abc

from ..common.model import Model


class Encoder(Model, abc.ABC):
    def __init__(self, num_dims, num_hidden, name='encoder', track_name='all'):
        super().__init__(name=name)
        self._track_name = track_name
        self._num_dims = num_dims
        self._num_hidden = [] if num_hidden is None else ([num_hidden] if isinstance(num_hidden, int)
                                                          else list(num_hidden))

    num_dims = property(lambda s: s._num_dims)
    num_hidden = property(lambda s: s._num_hidden)
    track_name = property(lambda s: s._track_name)

    @property
    def num_outputs(self):
        """Size of the encoding fed to the generators."""
        return self._num_hidden[-1] if self._num_hidden else self._num_dims

    @abc.abstractmethod
    def encode(self, x, u=None):
        """x[N,num_dims] -> (p_h, h)."""

    @abc.abstractmethod
    def decode(self, h, u=None):
        """h[N,num_outputs] -> (p_v, v)."""
