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

    def _layer_input(self, x, layer, u=None, seed=None):
        """Inputs of RBM `layer`: x through the frozen lower layers, SAMPLED (`self.encodings[layer - 1]`,
        dbn_encoder.py:213)."""
        h = x
        for i in range(layer):
            _, h = self._dbn.rbm_layers[i].forward(h, u=None if u is None else u['lower'][i], seed=seed)
        return h

    def init_bias(self, x, layer=0, u=None, seed=None):
        """The init_ops of dbn_encoder.py:237 / rbm.py:286-297: visible bias of RBM `layer` from the mean of its inputs."""
        assert 0 <= layer < self._dbn.num_layers
        self._dbn.rbm_layers[layer].visible_bias_init(self._layer_input(x, layer, u, seed))

    def layer_metrics(self, x, layer=0, u=None, seed=None):
        """The metrics dbn_encoder.py:219-231 builds for the trained RBM: one sampled up-down pass of its inputs,
        `batch/loss` = mean free-energy cost F(v) - F(v') (rbm.py:119, internal biases) and `log_likelihood` = mean over
        rows of the summed log-loss (tf.losses.log_loss, eps 1e-7, rbm.py:124-129). u = dict(lower=[..], up=, down=)."""
        import torch
        rbm = self._dbn.rbm_layers[layer]
        v = self._layer_input(x, layer, u, seed)
        _, h = rbm.forward(v, u=None if u is None else u['up'], seed=seed)
        p_v, v_s = rbm.reconstruct(h, u=None if u is None else u['down'], seed=seed)
        cost = rbm.free_energy_cost(v, v_s)
        cost = cost[0] if isinstance(cost, tuple) else cost
        eps = 1e-7
        ll = -(v * torch.log(p_v + eps) + (1 - v) * torch.log(1 - p_v + eps)).sum(1).mean()
        return {'batch/loss': cost, 'log_likelihood': ll}

    def train(self, x, lr, layer=0, u=None, seed=None):
        """Layer-wise CD-k (dbn_encoder.py:192-240): feed x through the frozen lower layers (sampled), then one
        RBM.train update on layer `layer`. u = dict(lower=[...], cd={...}) of uniforms or None."""
        assert 0 <= layer < self._dbn.num_layers                  # dbn_encoder.py:208
        h = x
        for i in range(layer):
            _, h = self._dbn.rbm_layers[i].forward(h, u=None if u is None else u['lower'][i], seed=seed)
        return self._dbn.rbm_layers[layer].train(h, lr, u=None if u is None else u['cd'], seed=seed)
