"""Base class of every module (mirrors reference models/common/model.py:9-234): a name, a build mode and
the trainable parameters. The TF graph/session of the reference is replaced by eager launches of the
sm_100a kernels; `build()` only validates the mode like model.py:115-149."""

BUILD_MODES = ('train', 'eval', 'generate')


class Model:
    def __init__(self, name='model'):
        self._name = name
        self._is_built = False
        self._mode = None
        self._metrics = {}

    @property
    def name(self):
        return self._name

    @property
    def is_built(self):
        return self._is_built

    @property
    def metrics(self):
        """dict of the last computed metrics (device tensors); `batch/loss` like metrics/statistical.py:34."""
        return self._metrics

    @property
    def trainable_params(self):
        return []

    @property
    def trainable_variables(self):
        return [p.data for p in self.trainable_params]

    def build(self, x=None, y=None, lengths=None, is_train=None, mode='eval'):
        if mode not in BUILD_MODES:
            raise ValueError("Incorrect build mode, supported modes are `train`, `eval`, and `generate`")
        self._mode = mode
        self._is_built = True
