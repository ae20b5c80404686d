"""Identity encoder (mirrors reference models/encoders/pass_encoder.py:13-136): encode/decode return their
input as both the probabilities and the samples (:99-127)."""
from .encoder import Encoder


class PassEncoder(Encoder):
    stochastic = False

    def __init__(self, num_dims, num_hidden=None, name='pass-encoder', track_name='all', arena=None):
        super().__init__(num_dims, None, name=name, track_name=track_name)

    def encode(self, x, u=None):
        return x, x

    def decode(self, h, u=None):
        return h, h
