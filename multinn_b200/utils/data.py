"""Dataset helpers (mirror reference multinn/utils/data.py:8-139; MIDI writing needs pypianoroll, which is absent, and is
out of scope). Pure NumPy: the arrays go to the device batch by batch in utils/training.py."""
import numpy as np


def pad_to_midi(songs, data_config):
    """data.py:8-33: [batch, time_steps, step_span, tracks] -> piano-rolls with 128 pitches [batch, steps, 128, tracks]."""
    pr = data_config['pitch_range']
    songs = np.reshape(songs, (songs.shape[0], -1, pr['highest'] - pr['lowest'], songs.shape[-1]))
    return np.pad(songs, ((0, 0), (0, 0), (pr['lowest'], 128 - pr['highest']), (0, 0)), 'constant', constant_values=0)


def reshape_songs(songs, step_size=1):
    """data.py:72-81: zero-pad time to a multiple of `step_size` pixels and fold them into the feature axis."""
    if songs.ndim != 4:
        raise ValueError("Dataset must have 4 dimensions.")
    pad_size = step_size - (songs.shape[1] % step_size)
    pad_size = 0 if pad_size == step_size else pad_size
    if pad_size > 0:
        songs = np.pad(songs, ((0, 0), (0, pad_size), (0, 0), (0, 0)), 'constant', constant_values=0)
    return songs.reshape([songs.shape[0], songs.shape[1] // step_size, songs.shape[2] * step_size, songs.shape[3]])


def load_data(data_config, step_size=1):
    """data.py:36-94: `<filename>.npy` piano-rolls [songs, time, pitches, tracks] (bool/uint8 as prepare_data.py stores
    them, or float32), optional lengths file, split into (train, valid, test) pairs of (data, lengths)."""
    path = data_config['filename']
    num_train = data_config['split']['num_train']
    num_valid = data_config['split']['num_valid']
    num_test = data_config['split']['num_test']
    if data_config['source'] != 'npy':
        raise ValueError('Not supported data format :(')
    songs = np.load(f'{path}.npy')[:num_train + num_valid + num_test]
    if len(songs.shape) != 4:
        raise ValueError("Dataset must have 4 dimensions.")
    if songs.shape[-1] != len(data_config['instruments']):
        raise ValueError(f"Dataset must have {len(data_config['instruments'])} tracks.")
    songs = reshape_songs(songs, step_size)
    if data_config.get('sequence_lengths'):
        lengths = np.load(data_config['sequence_lengths'])[:num_train + num_valid + num_test]
    else:
        lengths = np.full(songs.shape[0], songs.shape[1])
    train = (songs[:num_train], lengths[:num_train])
    valid = (songs[num_train:num_train + num_valid], lengths[num_train:num_train + num_valid])
    test = (songs[-num_test:], lengths[-num_test:])
    return train, valid, test


def prepare_sampling_inputs(X_train, X_valid, sampling_config, beat_size):
    """data.py:97-139: intro songs for sampling, the ids of the samples to save and their labels."""
    intro_steps = int(sampling_config['intro_beats'] * beat_size)
    intro_ids = sampling_config['intro_ids']
    intro_train = X_train[intro_ids['train']['start']:intro_ids['train']['end'], :intro_steps, :]
    intro_valid = X_valid[intro_ids['valid']['start']:intro_ids['valid']['end'], :intro_steps, :]
    intro_songs = np.concatenate([intro_train, intro_valid], axis=0)
    save_ids = sampling_config['save_ids']
    save_train = np.array(save_ids['train'])
    save_valid = np.array(save_ids['valid'])
    song_labels = [f't{i}' for i in save_train] + [f'v{i}' for i in save_valid]
    save_valid = save_valid + (intro_ids['train']['end'] - intro_ids['train']['start'])
    save_ids = np.concatenate([save_train, save_valid], axis=0)
    next_ids = save_ids
    for _ in range(1, sampling_config['num_save']):
        next_ids = next_ids + len(intro_songs)
        save_ids = np.concatenate([save_ids, next_ids], axis=0)
    return intro_songs, save_ids, song_labels
