"""Deep Belief Network = stack of RBMs (mirrors reference models/common/dbn.py:13-180)."""
from .model import Model
from .rbm import RBM


class DBN(Model):
    def __init__(self, num_dims, num_hidden, k=1, name='dbn', arena=None):
        super().__init__(name=name)
        if isinstance(num_hidden, int):
            num_hidden = [num_hidden]
        self._num_dims, self._num_hidden = num_dims, list(num_hidden)
        self.rbm_layers = []
        d = num_dims
        for i, h in enumerate(self._num_hidden):                      # dbn.py:44-54
            self.rbm_layers.append(RBM(d, h, k=k, name=f'{name}/rbm_{i}', arena=arena))
            d = h

    num_layers = property(lambda s: len(s.rbm_layers))
    num_dims = property(lambda s: s._num_dims)
    num_hidden = property(lambda s: s._num_hidden)

    @property
    def trainable_params(self):
        return [p for r in self.rbm_layers for p in r.trainable_params]

    def forward(self, x, u=None, seed=None):
        """dbn.py:136-156: chained sample-h through the layers. u = list of uniforms per layer. Returns (p_h, h)."""
        p, h = None, x
        for i, rbm in enumerate(self.rbm_layers):
            p, h = rbm.forward(h, u=None if u is None else u[i], seed=seed)
        return p, h

    def reconstruct(self, h, u=None, seed=None):
        """dbn.py:158-180: chained sample-v in reverse; u ordered like the loop (last layer first)."""
        p, v = None, h
        for j, rbm in enumerate(reversed(self.rbm_layers)):
            p, v = rbm.reconstruct(v, u=None if u is None else u[j], seed=seed)
        return p, v
