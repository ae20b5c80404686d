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
    """Loads the `.npy` piano-roll dataset the way reference utils/data.py:36-94 does and returns three (songs, lengths)
    pairs: train = the first `num_train` songs, valid = the next `num_valid`, test = the LAST `num_test` of the songs
    kept. Only the first num_train + num_valid + num_test songs of the file are used; `step_size` pixels are folded into
    the feature axis (reshape_songs); without a `sequence_lengths` file every song has the full (folded) length."""
    if data_config['source'] != 'npy':
        raise ValueError(f"data source {data_config['source']!r} is not supported (only 'npy')")
    counts = [data_config['split'][k] for k in ('num_train', 'num_valid', 'num_test')]
    used = sum(counts)
    songs = np.load(data_config['filename'] + '.npy')[:used]
    if songs.ndim == 4 and songs.shape[3] != len(data_config['instruments']):
        raise ValueError(f"the dataset has {songs.shape[3]} tracks, the config names {len(data_config['instruments'])}")
    songs = reshape_songs(songs, step_size)
    lengths_file = data_config.get('sequence_lengths')
    lengths = np.load(lengths_file)[:used] if lengths_file else np.full(len(songs), songs.shape[1])
    n_tr, n_va, n_te = counts
    cut = lambda sl: (songs[sl], lengths[sl])
    return cut(slice(0, n_tr)), cut(slice(n_tr, n_tr + n_va)), cut(slice(len(songs) - n_te, len(songs)))


def prepare_sampling_inputs(X_train, X_valid, sampling_config, beat_size):
    """Picks the intro songs of a sampling run (role of reference utils/data.py:97-138). Returns
    (intro_songs, save_ids, song_labels): the first `intro_beats` beats of the configured train and valid song ranges,
    concatenated; the rows of the TILED sample batch (generate_music repeats the intros `num_songs` times) that are
    written to disk -- `num_save` repeats of the chosen train ids and (shifted past the train intros) valid ids; and the
    labels 't<i>' / 'v<i>' of those songs."""
    steps = int(sampling_config['intro_beats'] * beat_size)
    ranges, chosen = sampling_config['intro_ids'], sampling_config['save_ids']
    take = lambda X, r: X[r['start']:r['end'], :steps]
    intro_songs = np.concatenate([take(X_train, ranges['train']), take(X_valid, ranges['valid'])], axis=0)
    n_train_intros = ranges['train']['end'] - ranges['train']['start']
    first = np.concatenate([np.asarray(chosen['train'], dtype=np.int64),
                            np.asarray(chosen['valid'], dtype=np.int64) + n_train_intros])
    repeats = np.arange(sampling_config['num_save'])[:, None] * len(intro_songs)
    save_ids = (first[None, :] + repeats).reshape(-1)
    song_labels = [f't{i}' for i in chosen['train']] + [f'v{i}' for i in chosen['valid']]
    return intro_songs, save_ids, song_labels


# ----------------------------------------------------------------------------- MIDI export (data.py:141-202)
# The reference hands the tracks to pypianoroll 0.5.0 (`Multitrack(...).write(path)`), which is not installed here and has
# no wheel in the image. write_song() below writes the Standard MIDI File itself: same per-track gains, programs, drum
# flags, tempo and beat resolution; a note is a maximal run of non-zero steps of one pitch, its velocity the value at
# the onset. Byte-level equality with pypianoroll's output is NOT claimed (it cannot be checked here).
TRACK_GAIN = {'Piano': 0.8, 'Strings': 0.9, 'Bass': 1.2}          # data.py:163-169 "manually adjusting the sound"


def _vlq(n):
    """MIDI variable-length quantity."""
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def _chunk(tag, body):
    return tag + len(body).to_bytes(4, 'big') + body


def track_notes(pianoroll):
    """[steps, 128] velocities -> list of (onset_step, end_step, pitch, velocity), end exclusive."""
    on = np.asarray(pianoroll) > 0
    edge = np.zeros((on.shape[0] + 2, on.shape[1]), np.int8)
    edge[1:-1] = on
    d = np.diff(edge, axis=0)
    p_on, t_on = np.nonzero(d.T > 0)
    _, t_off = np.nonzero(d.T < 0)
    vel = np.clip(np.rint(np.asarray(pianoroll)[t_on, p_on]), 1, 127).astype(int)
    return [(int(a), int(b), int(p), int(v)) for a, b, p, v in zip(t_on, t_off, p_on, vel)]


def song_to_midi_bytes(song, data_config):
    """song[time_steps, 128, tracks] in {0,1} -> bytes of a format-1 Standard MIDI File (one conductor track with the
    tempo + one track per instrument; ticks per quarter note = beat_resolution, so one step = one tick)."""
    song = np.asarray(song, np.float32) * 100.                     # data.py:155
    names = data_config['instruments']
    chunks = [_chunk(b'MTrk', _vlq(0) + b'\xff\x51\x03' + int(round(60e6 / data_config['tempo'])).to_bytes(3, 'big')
                     + _vlq(0) + b'\xff\x2f\x00')]
    melodic = [c for c in range(16) if c != 9]
    for i, name in enumerate(names):
        roll = song[..., i] * TRACK_GAIN.get(name, 1.0)
        drum = bool(data_config['is_drums'][i])
        ch = 9 if drum else melodic[i % len(melodic)]
        events = []                                                # (tick, order, bytes): note-offs sort before note-ons
        for t0, t1, pitch, vel in track_notes(roll):
            events.append((t0, 1, bytes([0x90 | ch, pitch, vel])))
            events.append((t1, 0, bytes([0x80 | ch, pitch, 0])))
        events.sort(key=lambda e: (e[0], e[1]))
        body = _vlq(0) + b'\xff\x03' + _vlq(len(name.encode())) + name.encode()
        body += _vlq(0) + bytes([0xC0 | ch, int(data_config['programs'][i]) & 0x7F])
        now = 0
        for tick, _, msg in events:
            body += _vlq(tick - now) + msg
            now = tick
        chunks.append(_chunk(b'MTrk', body + _vlq(0) + b'\xff\x2f\x00'))
    header = _chunk(b'MThd', (1).to_bytes(2, 'big') + len(chunks).to_bytes(2, 'big')
                    + int(data_config['beat_resolution']).to_bytes(2, 'big'))
    return header + b''.join(chunks)


def write_song(song, path, data_config):
    """data.py:141-178: one multi-track piano-roll song [time_steps, 128, tracks] -> MIDI file."""
    with open(path, 'wb') as f:
        f.write(song_to_midi_bytes(song, data_config))


def save_music(music, num_intro, data_config, base_path, save_dir='outputs/', song_labels=None):
    """data.py:181-202: music[num_songs * num_intro, time_steps, 128, tracks]; file names as in the reference."""
    import os
    os.makedirs(save_dir, exist_ok=True)
    paths = []
    for i in range(num_intro):
        for j in range(music.shape[0] // num_intro):
            label = f'song{i}' if song_labels is None else f'{song_labels[i]}'
            paths.append(os.path.join(save_dir, f'{base_path}_{label}_{j}.mid'))
            write_song(music[i + j * num_intro], paths[-1], data_config)
    return paths
