"""Musical metrics of generated piano-rolls (mirrors reference multinn/metrics/musical.py:45-275 and the
summary layout of metrics/musical_tf.py:158-229). Host-side NumPy reporting code, not on the training path: samples
come back from the device once per sampling run (sample.py:95-115).

Bar music is `[batch, bars, 4 * beat_resolution, pitch_span, tracks]`, binary (bool or {0,1} numbers).
Pinned against the reference's own NumPy module by tests/golden/musical_metrics.npz (tools/make_golden_musical.py runs
/root/reference/multinn/metrics/musical.py on seeded rolls). Reference quirks kept on purpose:
  M1  chroma folds CONSECUTIVE pitches: class c = pitches [c*P/12, (c+1)*P/12), not p % 12 (musical.py:37-42);
  M2  the qualified-note denominator skips an onset at flat position 0 of a track, i.e. a note of (sample 0, lowest
      pitch) that sounds on the very first step (`count_nonzero` of onset POSITIONS, musical.py:107-110);
  M3  qualified means strictly longer than `threshold` steps, polyphonic strictly more than `threshold` pitches.
"""
import numpy as np

TRACK_HEADER = ('Drums', 'Piano', 'Guitar', 'Bass', 'Strings')


def _need_dims(a, n):
    if np.ndim(a) != n:
        raise ValueError(f"Input tensor must have {n} dimensions.")


def to_bars(x, beat_resolution, pitch_span):
    """core/multinn_core.py:354-358: model-format music [batch, time_steps, num_dims, tracks] -> bar music
    [batch, bars, 4 * beat_resolution, pitch_span, tracks]."""
    x = np.asarray(x)
    return x.reshape(x.shape[0], -1, 4 * beat_resolution, pitch_span, x.shape[-1])


def to_chroma(pianoroll):
    """musical.py:15-42 (quirk M1). [..., pitch, tracks] -> [..., 12, tracks] note counts per class."""
    roll = np.asarray(pianoroll)
    P = roll.shape[-2]
    short = (-P) % 12
    if short:
        widths = [(0, 0)] * roll.ndim
        widths[-2] = (0, short)
        roll = np.pad(roll, widths, 'constant')
        P += short
    per_class = P // 12
    folded = roll.reshape(roll.shape[:-2] + (12, per_class, roll.shape[-1]))
    return folded.sum(axis=-2)


def empty_bar_rate(pianoroll):
    """musical.py:45-58: share of (sample, bar) cells without any note, per track."""
    _need_dims(pianoroll, 5)
    roll = np.asarray(pianoroll)
    has_note = roll.astype(bool).reshape(roll.shape[0] * roll.shape[1], -1, roll.shape[-1]).any(axis=1)
    return 1 - has_note.mean(axis=0)


def num_pitches_used(pianoroll):
    """musical.py:61-74: mean number of distinct pitches (or pitch classes) sounding in a bar, per track."""
    _need_dims(pianoroll, 5)
    used = np.asarray(pianoroll).astype(bool).any(axis=2)            # [batch, bars, pitch, tracks]
    return used.sum(axis=2).mean(axis=(0, 1))


def _note_runs(track_roll):
    """Onset and end positions of every maximal run of ones along time. track_roll[samples, steps, pitch] ->
    (onset_step, end_step, sample, pitch) arrays; end is exclusive."""
    on = np.asarray(track_roll).astype(bool).transpose(0, 2, 1)       # [samples, pitch, steps]
    S, P, T = on.shape
    edge = np.zeros((S, P, T + 2), np.int8)
    edge[:, :, 1:-1] = on
    step = np.diff(edge, axis=2)                                      # +1 at an onset, -1 one past the last step
    s_on, p_on, t_on = np.nonzero(step > 0)
    _, _, t_off = np.nonzero(step < 0)                                # same (sample, pitch)-major order: pairs line up
    return t_on, t_off, s_on, p_on


def qualified_note_rate(pianoroll, threshold=2):
    """musical.py:77-113: notes lasting more than `threshold` steps (bars of a sample joined) over the onset count
    (quirks M2, M3). nan for a track without counted onsets (0/0), inf if only the skipped onset exists."""
    _need_dims(pianoroll, 5)
    roll = np.asarray(pianoroll)
    B, bars, steps, P, M = roll.shape
    joined = roll.reshape(B, bars * steps, P, M)
    out = np.empty(M, np.float32)
    for m in range(M):
        t_on, t_off, s_on, p_on = _note_runs(joined[..., m])
        qualified = np.float32(np.count_nonzero(t_off - t_on > threshold))
        counted = np.float32(t_on.size - np.count_nonzero((s_on == 0) & (p_on == 0) & (t_on == 0)))   # M2
        with np.errstate(divide='ignore', invalid='ignore'):
            out[m] = qualified / counted
    return out


def polyphonic_rate(pianoroll, threshold=2):
    """musical.py:116-132: share of time steps with more than `threshold` simultaneous pitches, per track."""
    _need_dims(pianoroll, 5)
    roll = np.asarray(pianoroll)
    crowded = np.count_nonzero(roll, axis=3) > threshold              # [batch, bars, steps, tracks]
    return (crowded.sum(axis=2) / roll.shape[2]).mean(axis=(0, 1))


_DRUM_CELLS = {   # steps per bar -> (pattern of one cell, repeats); tolerance fills the `t` entries (musical.py:146-165)
    96: ((1, 't', 0, 0, 0, 't'), 16), 48: ((1, 't', 't'), 16), 24: ((1, 't', 't'), 8),
    72: ((1, 't', 0, 0, 0, 't'), 12), 36: ((1, 't', 't'), 12), 64: ((1, 't', 0, 't'), 16),
    32: ((1, 't'), 16), 16: ((1, 't'), 8),
}


def drum_pattern_mask(n_timesteps, tolerance=0.1):
    if n_timesteps not in _DRUM_CELLS:
        raise ValueError("Unsupported number of timesteps for the drum in pattern metric.")
    cell, reps = _DRUM_CELLS[n_timesteps]
    return np.tile([tolerance if c == 't' else float(c) for c in cell], reps)


def drum_in_pattern_rate(drums):
    """musical.py:135-173: drums[batch, bars, steps, pitch-or-class] -> weighted share of drum hits on the beat grid."""
    _need_dims(drums, 4)
    d = np.asarray(drums)
    hits_per_step = d.sum(axis=3)
    weighted = float((hits_per_step * drum_pattern_mask(d.shape[2])[None, None, :]).sum())
    total = np.count_nonzero(d)
    return weighted / total if total > 0 else 0.


def tonal_matrix(r1=1.0, r2=1.0, r3=0.5):
    """Harte et al. 2006 6-D tonal centroid transform (musical.py:196-209): circles of fifths, minor and major thirds."""
    pc = np.arange(12)
    rows = []
    for radius, turn in ((r1, 7. / 6.), (r2, 3. / 2.), (r3, 2. / 3.)):
        rows += [radius * np.sin(pc * turn * np.pi), radius * np.cos(pc * turn * np.pi)]
    return np.stack(rows)


def harmonicity(chroma):
    """musical.py:176-231: mean tonal distance between every pair of tracks. chroma[batch, bars, steps, 12, tracks] ->
    [tracks, tracks]; beats (a quarter of a bar) in which either track is silent are left out of the mean (nan)."""
    _need_dims(chroma, 5)
    c = np.asarray(chroma)
    if c.shape[3] != 12:
        raise ValueError("Input tensor must be a chroma tensor.")
    M = c.shape[4]
    per_beat = c.reshape(-1, c.shape[2] // 4, 12, M).sum(axis=1)                 # [beats, 12, tracks]
    with np.errstate(divide='ignore', invalid='ignore'):
        per_beat = per_beat / per_beat.sum(axis=1, keepdims=True)
        flat = per_beat.transpose(1, 0, 2).reshape(12, -1)
        centroid = (tonal_matrix() @ flat).reshape(6, -1, M)                     # [6, beats, tracks]
        gap = centroid[:, :, :, None] - centroid[:, :, None, :]
        dist = np.sqrt((np.abs(gap) ** 2).sum(axis=0))                           # [beats, tracks, tracks]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)                      # all-nan pairs stay nan
            return np.nanmean(dist, axis=0)


# ----------------------------------------------------------------------------- evaluation entry points
def sample_metrics(pianoroll):
    """The numbers compute_sample_metrics prints (musical.py:234-275), as a dict: track 0 is the drum track; UPC, QN, PR
    and TD cover the pitched tracks 1.."""
    _need_dims(pianoroll, 5)
    roll = np.asarray(pianoroll)
    chroma = to_chroma(roll[..., 1:])
    return {
        'EB': empty_bar_rate(roll), 'UP': num_pitches_used(roll), 'UPC': num_pitches_used(chroma),
        'QN': np.array([qualified_note_rate(roll[..., i:i + 1])[0] for i in range(1, roll.shape[-1])], np.float32),
        'PR': polyphonic_rate(roll[..., 1:]), 'DP': drum_in_pattern_rate(roll[..., 0]), 'TD': harmonicity(chroma),
    }


def format_sample_metrics(metrics, header=TRACK_HEADER):
    """The console table of compute_sample_metrics as a string."""
    cell = lambda vals: '  '.join(f'{v:.5f}' for v in vals)
    blank = '   -     '
    lines = ['', ' ' * 5 + ' ' + '    '.join(header),
             f'{"EB: ":5s}' + cell(metrics['EB']), f'{"UP: ":5s}' + cell(metrics['UP']),
             f'{"UPC: ":5s}' + blank + cell(metrics['UPC']), f'{"QN: ":5s}' + blank + cell(metrics['QN']),
             f'{"PR: ":5s}' + blank + cell(metrics['PR']), '', f'{"DP: ":5s}{metrics["DP"]:.5f}', '', f'{"TD: ":5s}',
             str(metrics['TD'])]
    return '\n'.join(lines)


def metric_ops(bar_music):
    """musical_tf.py:158-183 (`get_metric_ops`): every metric over ALL tracks of the bar music, DP on track 0 and TD on
    the chroma of tracks 1.. (TF versions threshold at 0.5, identical on binary rolls)."""
    _need_dims(bar_music, 5)
    roll = np.asarray(bar_music) > 0.5
    chroma = to_chroma(roll)
    return {'EB': empty_bar_rate(roll), 'UP': num_pitches_used(roll), 'UPC': num_pitches_used(chroma),
            'PR': polyphonic_rate(roll), 'QN': qualified_note_rate(roll), 'DP': drum_in_pattern_rate(roll[..., 0]),
            'TD': harmonicity(chroma[..., 1:])}


def metric_summary(bar_music, tracks):
    """musical_tf.py:186-229 (`get_metric_summary_ops`): the scalars the reference writes to TensorBoard, keyed by their
    summary scope: intra-track EB/UP (+DP for Drums, +UPC/QN/PR otherwise), inter-track tonal distances."""
    ops = metric_ops(bar_music)
    out = {}
    for i, track in enumerate(tracks):
        scope = f'sample_scores/intra-track/{track}'
        out[f'{scope}/EB'] = float(ops['EB'][i])
        out[f'{scope}/UP'] = float(ops['UP'][i])
        if track == 'Drums':
            out[f'{scope}/DP'] = float(ops['DP'])
        else:
            for name in ('UPC', 'QN', 'PR'):
                out[f'{scope}/{name}'] = float(ops[name][i])
    for i in range(1, len(tracks)):
        for j in range(i + 1, len(tracks)):
            pair = f'{tracks[i]}-{tracks[j]}'          # scope TD/<a>-<b> AND a scalar of the same name (musical_tf.py:222-225)
            out[f'sample_scores/inter-track/TD/{pair}/{pair}'] = float(ops['TD'][i - 1][j - 1])
    return out
