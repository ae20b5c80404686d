"""RNN-MultiNADE generator (mirrors reference models/generators/rnn_multinade.py:15-317): one shared temporal
LSTM, a Dense layer of M*(D+H) units and M NADEs, one per track. Used by the Composer mode."""
from .rnn_nade import RnnNade


class RnnMultiNADE(RnnNade):
    def __init__(self, num_dims, num_hidden, num_hidden_rnn, tracks, keep_prob=1.0, internal_bias=False,
                 name='rnn-multinade', arena=None, num_inputs=None):
        self._tracks = list(tracks)
        super().__init__(num_dims=num_dims, num_hidden=num_hidden, num_hidden_rnn=num_hidden_rnn,
                         keep_prob=keep_prob, internal_bias=internal_bias, name=name, track_name='all',
                         arena=arena, num_inputs=num_inputs, num_tracks=len(self._tracks))
        self._num_output = self.num_tracks * self.num_dims

    @property
    def tracks(self):
        return self._tracks
