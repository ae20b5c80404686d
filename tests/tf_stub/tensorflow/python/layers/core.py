"""TEST INFRASTRUCTURE ONLY -- `tensorflow.python.layers.core.Dense` of the NumPy TF stand-in: y = activation(x . kernel + bias),
kernel [in, units] created on first call with the given initializer (glorot_uniform by default, as tf.layers.Dense), bias
zeros (generators/rnn_nade.py:54-57, rnn_multinade.py:60-63, common/dnn.py:56-60)."""
import numpy as np

import tensorflow as tf


class Dense:
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None, name='dense',
                 **kw):
        self.units, self.activation, self.use_bias, self.name = units, activation, use_bias, name
        self._kinit = kernel_initializer or tf.contrib.layers.xavier_initializer()
        self.kernel = self.bias = None

    def build(self, in_dim):
        # tf.layers.Layer._set_scope: variable_scope(None, default_name='dense') -> 'dense', 'dense_1', ... unique
        # within the enclosing scope, decided at the first call
        self.name = tf.unique_default_name(self.name)
        with tf.variable_scope(self.name):
            self.kernel = tf.Variable(self._kinit([int(in_dim), self.units]), name='kernel')
            self.bias = tf.Variable(np.zeros(self.units), name='bias') if self.use_bias else None

    @property
    def trainable_variables(self):
        return [v for v in (self.kernel, self.bias) if v is not None]

    variables = trainable_variables

    def __call__(self, inputs):
        x = np.asarray(inputs)
        if self.kernel is None:
            self.build(x.shape[-1])
        y = x @ np.asarray(self.kernel)
        if self.use_bias:
            y = y + np.asarray(self.bias)
        y = tf._t(y)
        return self.activation(y) if self.activation is not None else y

    apply = __call__
