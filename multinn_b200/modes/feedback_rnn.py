"""Feedback-RNN MultINN (mirrors reference models/multinn/multinn_feedback_rnn.py:14-79): the feedback module is an LSTM
stack run (tf.nn.dynamic_rnn) over the stacked encodings of all T+1 padded steps; during generation its state is carried
from the final state over the padded intro (multinn_feedback.py:79-82,159)."""
from ..common.rnn import RNN
from .feedback import MultINNFeedback


class MultINNFeedbackRnn(MultINNFeedback):
    def __init__(self, config, params, name='MultINN-feedback-rnn', **kw):
        super().__init__(config, params, name=name, **kw)
        self._mode = 'feedback-rnn'

    def _init_feedback(self):
        """multinn_feedback_rnn.py:34-39: RNN(num_units=feedback, keep_prob)."""
        return RNN(self._arena, self._num_dims_generator * self.num_tracks, self._feedback_units(),
                   keep_prob=self.keep_prob, name='feedback/rnn', binary_inputs=True)

    def _apply_feedback(self, stack, keep=1.0, u_fb=None, seed=0, save=True):
        out, state = self._feedback_layer.forward_sequence(stack.contiguous(), keep=keep, u=u_fb, seed=seed)
        self._fb_final_state = [type(s)(s[0].clone(), s[1].clone()) for s in state]
        return out

    def _feedback_backward(self, dfb):
        self._feedback_layer.backward_sequence(dfb)

    def _feedback_intro_state(self):
        return self._fb_final_state

    def _feedback_step(self, samples_stack, state):
        return self._feedback_layer.step(samples_stack, state)
