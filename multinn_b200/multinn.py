"""MultINN facade (mirrors reference models/multinn/multinn.py:9-299): picks the operation mode."""
from .modes.composer import MultINNComposer
from .modes.jamming import MultINNJamming


def _lazy(modname, clsname):
    def make(*a, **kw):
        import importlib
        return getattr(importlib.import_module(f'multinn_b200.modes.{modname}'), clsname)(*a, **kw)
    return make


_MODES = {
    'composer': MultINNComposer,
    'jamming': MultINNJamming,
    'joint': _lazy('joint', 'MultINNJoint'),
    'feedback': _lazy('feedback', 'MultINNFeedback'),
    'feedback-rnn': _lazy('feedback_rnn', 'MultINNFeedbackRnn'),
}


class MultINN:
    def __init__(self, config, params, mode='feedback-rnn', name='MultINN', **kw):
        if mode not in _MODES:
            raise ValueError('Incorrect operation mode, choose from `joint`, `composer`, `jamming`, `feedback`, '
                             'and `feedback-rnn`.')            # multinn.py:37-49
        self._model = _MODES[mode](config, params, name=name, **kw)

    def __getattr__(self, item):
        return getattr(self.__dict__['_model'], item)


def default_config(num_pixels=1, instruments=('Drums', 'Piano', 'Guitar', 'Bass', 'Strings'), beat_resolution=12):
    """The keys of reference configs/default_config.yaml that the hot path reads."""
    return {'data': {'beat_resolution': beat_resolution, 'pitch_range': {'lowest': 24, 'highest': 108},
                     'instruments': list(instruments), 'programs': [0, 0, 24, 32, 48][:len(instruments)],
                     'is_drums': [True, False, False, False, False][:len(instruments)], 'tempo': 120},
            'training': {'random_seed': 23, 'batch_size': 32, 'num_pixels': num_pixels, 'piece_size': 16,
                         'learning_rate': 0.01, 'clip_norm': 5.},
            'sampling': {'num_songs': 3, 'intro_beats': 8, 'sample_beats': 88}}


def default_params(mode='composer', encoder='Pass', encoder_hidden=None, generator='NADE', num_hidden=256,
                   num_hidden_rnn=(512, 256), feedback=None, keep_prob=0.9, tune_encoder=False):
    """The keys of reference configs/default_params.yaml."""
    return {'mode': mode, 'tune_encoder': tune_encoder, 'keep_prob': keep_prob,
            'encoder': {'type': encoder, 'num_hidden': encoder_hidden},
            'generator': {'type': generator, 'num_hidden': num_hidden, 'num_hidden_rnn': list(num_hidden_rnn),
                          'feedback': feedback}}
