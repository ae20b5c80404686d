"""Multi-layer sigmoid Dense network (mirrors reference models/common/dnn.py:13-138), the Feedback module of the
Feedback MultINN: every layer is tf.layers.Dense(units, activation=sigmoid, xavier init) (dnn.py:56-60)."""
import torch

from .. import ops
from ..params import glorot_uniform, zeros
from .model import Model


class DNN(Model):
    def __init__(self, arena, num_inputs, num_units=128, name='dnn'):
        super().__init__(name=name)
        if isinstance(num_units, int):
            num_units = [num_units]
        self._num_units = list(num_units)
        self.kernels, self.biases = [], []
        i = num_inputs
        for l, u in enumerate(self._num_units):
            self.kernels.append(arena.add(f'{name}/dense_{l}/kernel', (i, u), glorot_uniform(i, u)))
            self.biases.append(arena.add(f'{name}/dense_{l}/bias', (u,), zeros()))
            i = u
        self._saved = None

    num_units = property(lambda s: s._num_units)
    num_layers = property(lambda s: len(s._num_units))

    @property
    def trainable_params(self):
        return [p for pair in zip(self.kernels, self.biases) for p in pair]

    def __call__(self, inputs, save=True):
        """dnn.py:97-116: x -> sigmoid(x K + b) per layer. inputs[N,in] -> [N,num_units[-1]]."""
        acts = [inputs]
        x = inputs
        for k, b in zip(self.kernels, self.biases):
            pre = torch.empty(x.shape[0], k.shape[1], device=x.device)
            ops.gemm(x, k.data, pre)
            ops.bias_sigmoid_sample(pre, bias=b.data.view(1, -1), p=pre)       # in place: pre <- sigmoid(pre + b)
            acts.append(pre)
            x = pre
        if save:
            self._saved = acts
        return x

    def backward(self, dy):
        """Gradients of the layers' kernels/biases given d(loss)/d(outputs); the inputs are stop-gradient encodings."""
        acts = self._saved
        for l in reversed(range(self.num_layers)):
            y, x = acts[l + 1], acts[l]
            dpre = torch.empty_like(y)
            ops.sigmoid_bwd(y, dy, dpre)
            ops.gemm(x, dpre, self.kernels[l].grad, transA=True)
            ops.colsum(dpre, self.biases[l].grad)
            if l > 0:
                dy = torch.empty_like(x)
                ops.gemm(dpre, self.kernels[l].data, dy, transB=True)
