"""DBN encoder (mirrors reference models/encoders/dbn_encoder.py:15-240): per-(b,t)-row DBN; `encode()` returns the
SAMPLED binary last-layer code (quirk Q12: stochastic, re-sampled on every call), `decode()` the sampled visible."""
from ..common.dbn import DBN
from .encoder import Encoder


class DBNEncoder(Encoder):
    stochastic = True

    def __init__(self, num_dims, num_hidden, name='dbn-encoder', track_name='all', arena=None, k=1):
        super().__init__(num_dims, num_hidden, name=name, track_name=track_name)
        if not self._num_hidden:
            raise ValueError('DBNEncoder needs `encoder.num_hidden`, e.g. [168, 84]')
        self._dbn = DBN(num_dims, self._num_hidden, k=k, name=name, arena=arena)

    dbn = property(lambda s: s._dbn)

    @property
    def trainable_params(self):
        return self._dbn.trainable_params

    def encode(self, x, u=None, seed=None):
        """x[N,num_dims] -> (p_h, h) of the last layer (dbn_encoder.py:136-162)."""
        return self._dbn.forward(x, u=u, seed=seed)

    def decode(self, h, u=None, seed=None):
        """h[N,num_hidden[-1]] -> (p_v, v) on the input layer (dbn_encoder.py:164-190)."""
        return self._dbn.reconstruct(h, u=u, seed=seed)

    def train(self, x, lr, layer=0, u=None, seed=None):
        """Layer-wise CD-k (dbn_encoder.py:192-240): feed x through the frozen lower layers (sampled), then one
        RBM.train update on layer `layer`. u = dict(lower=[...], cd={...}) of uniforms or None."""
        assert 0 <= layer < self._dbn.num_layers                  # dbn_encoder.py:208
        h = x
        for i in range(layer):
            _, h = self._dbn.rbm_layers[i].forward(h, u=None if u is None else u['lower'][i], seed=seed)
        return self._dbn.rbm_layers[layer].train(h, lr, u=None if u is None else u['cd'], seed=seed)
