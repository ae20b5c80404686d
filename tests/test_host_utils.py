"""Host-side utilities around the hot path (row f1 of the scope table): dataset reshaping / splitting, the batch pieces
of train.py:165-178 and utils/training.py:201-213, TrainingStats / LossAccumulator, and the epoch loop with early
stopping -- checked on CPU with a stub model (the GPU run of the same loop is in test_gpu_model.py)."""
import numpy as np
import pytest

from multinn_b200.utils import data as D
from multinn_b200.utils import training as U


def test_training_pieces_follow_the_reference_loop():
    rng = np.random.default_rng(0)
    X = rng.random((5, 10, 3, 2)) < 0.5
    lengths = np.array([10, 3, 7, 1, 10])
    ids = np.array([4, 1, 0, 3, 2])
    got = list(U.training_pieces(X, lengths, ids, batch_size=2, piece_size=4))
    # literal restatement of train.py:165-178
    ref = []
    for bi, i in enumerate(range(0, 5, 2)):
        for j in range(0, 10, 4):
            lb = lengths[ids[i:i + 2]] - j
            ne = np.where(lb > 0)[0]
            if len(ne) > 0:
                lb = np.minimum(lb[ne], 4)
                ref.append((bi, X[ids[i:i + 2], j:j + lb.max(), ...][ne], lb))
    assert len(got) == len(ref) == 8
    for (b0, s0, l0), (b1, s1, l1) in zip(got, ref):
        assert b0 == b1 and np.array_equal(s0, s1) and np.array_equal(l0, l1)
        assert s0.shape[1] == l0.max() and (l0 > 0).all()


def test_evaluation_pieces_keep_quirk_q9():
    X = np.zeros((3, 8, 2, 1), dtype=np.float32)
    lengths = np.array([8, 2, 5])
    pieces = list(U.evaluation_pieces(X, lengths, batch_size=4, piece_size=4))
    assert len(pieces) == 2
    for songs, seq in pieces:                       # the offset j is never subtracted from the lengths
        assert songs.shape == (3, 4, 2, 1) and np.array_equal(seq, [4, 2, 4])


def test_training_stats_and_loss_accumulator(tmp_path):
    st = U.TrainingStats()
    st.new_run(); st.new_epoch(); st.new_step(); st.new_step(); st.update_metric_best(12.5); st.new_idle_epoch()
    st.save(str(tmp_path / 'steps'))
    other = U.TrainingStats()
    other.load(str(tmp_path / 'steps'))
    assert (other.steps, other.epoch, other.run, other.metric_best, other.idle_epochs) == (2, 1, 1, 12.5, 0)
    acc = U.LossAccumulator()
    for v in (1.0, float('nan'), 3.0, float('inf'), float('-inf')):
        acc.update(v)
    assert acc.loss() == 2.0 and acc.num_bad() == 3 and abs(acc.ratio_bad() - 0.6) < 1e-12
    assert 'nan: 1' in str(acc) and '+inf: 1' in str(acc)
    acc.clear()
    assert np.isnan(acc.loss())


def test_load_data_split_reshape_and_errors(tmp_path):
    rng = np.random.default_rng(1)
    songs = rng.random((7, 9, 4, 2)) < 0.3
    np.save(tmp_path / 'songs.npy', songs)
    np.save(tmp_path / 'lengths.npy', np.array([9, 8, 7, 6, 5, 4, 3]))
    cfg = {'filename': str(tmp_path / 'songs'), 'source': 'npy', 'instruments': ['a', 'b'],
           'sequence_lengths': str(tmp_path / 'lengths.npy'), 'split': {'num_train': 4, 'num_valid': 2, 'num_test': 1}}
    (xt, lt), (xv, lv), (xs, ls) = D.load_data(cfg, step_size=2)
    assert xt.shape == (4, 5, 8, 2) and xv.shape == (2, 5, 8, 2) and xs.shape == (1, 5, 8, 2)
    assert np.array_equal(lt, [9, 8, 7, 6]) and np.array_equal(lv, [5, 4]) and np.array_equal(ls, [3])
    # pixels folded into the feature axis: step t, feature p*4 + d = time 2t + p, pitch d; one zero step appended
    assert np.array_equal(xt[1, 2, :4, 0], songs[1, 4, :, 0]) and np.array_equal(xt[1, 2, 4:, 1], songs[1, 5, :, 1])
    assert not xt[:, 4, 4:].any()
    cfg2 = dict(cfg, sequence_lengths=None)
    (xt2, lt2), _, _ = D.load_data(cfg2, step_size=1)
    assert xt2.shape == (4, 9, 4, 2) and np.array_equal(lt2, [9, 9, 9, 9])
    with pytest.raises(ValueError):
        D.load_data(dict(cfg, instruments=['a']), 1)
    with pytest.raises(ValueError):
        D.load_data(dict(cfg, source='tfrecord'), 1)


def test_pad_to_midi_and_sampling_inputs():
    cfg = {'pitch_range': {'lowest': 24, 'highest': 108}}
    x = np.ones((2, 3, 168, 5), dtype=np.float32)              # num_pixels = 2
    p = D.pad_to_midi(x, cfg)
    assert p.shape == (2, 6, 128, 5) and p[:, :, :24].sum() == 0 and p[:, :, 108:].sum() == 0 and p[:, :, 24:108].all()
    Xt = np.arange(6 * 10).reshape(6, 10, 1, 1).astype(np.float32)
    Xv = -np.arange(4 * 10).reshape(4, 10, 1, 1).astype(np.float32)
    sc = {'intro_beats': 2, 'num_save': 2, 'intro_ids': {'train': {'start': 1, 'end': 4}, 'valid': {'start': 0, 'end': 2}},
          'save_ids': {'train': [0, 2], 'valid': [1]}}
    intro, save_ids, labels = D.prepare_sampling_inputs(Xt, Xv, sc, beat_size=3.0)
    assert intro.shape == (5, 6, 1, 1) and labels == ['t0', 't2', 'v1']
    assert np.array_equal(save_ids, [0, 2, 4, 5, 7, 9])


class _StubModel:
    """Loss falls for three epochs, then the validation metric stalls."""

    def __init__(self):
        self.calls, self.saved, self.valid = [], [], iter([5.0, 4.0, 4.5, 4.2, 4.1, 3.0])
        self.epoch_metric = None

    def train_generators(self, optimizer, lr):
        import torch

        def step(x, lengths=None, seed=None):
            self.calls.append((tuple(x.shape), lengths.tolist(), seed))
            return torch.tensor([float(len(self.calls))])
        return step

    def evaluate(self, x, lengths=None):
        import torch
        if self.epoch_metric is None:
            self.epoch_metric = next(self.valid)
        return {'nll': torch.full((int(lengths.sum()), 2), self.epoch_metric)}

    def save(self, path):
        self.saved.append(path)


def test_fit_early_stopping_checkpoints_and_step_counter(tmp_path):
    X = (np.random.default_rng(2).random((5, 8, 3, 2)) < 0.3).astype(np.uint8)
    lengths = np.array([8, 8, 3, 5, 8])
    model = _StubModel()
    orig = U.collect_metrics

    def collect(model_, *a, **kw):
        model_.epoch_metric = None
        return orig(model_, *a, **kw)
    U_collect, U.collect_metrics = U.collect_metrics, collect
    try:
        cfg = {'batch_size': 2, 'piece_size': 4, 'learning_rate': 0.01, 'epochs': 6, 'early_stopping': 2}
        stats, hist = U.fit(model, (X, lengths), (X[:2], lengths[:2]), cfg, checkpoint_path=str(tmp_path / 'best.pt'),
                            device='cpu')
    finally:
        U.collect_metrics = U_collect
    # epochs 1, 2 improve (5.0, 4.0); 3 and 4 do not -> stop after epoch 4
    assert [h['epoch'] for h in hist] == [1, 2, 3, 4]
    assert stats.metric_best == 4.0 and stats.idle_epochs == 2 and len(model.saved) == 2
    assert stats.steps == 4 * 3 and stats.run == 1            # 3 batches per epoch, one step per batch (Q11)
    assert (tmp_path / 'best.pt.stats').exists()
    # every fed piece: at most piece_size frames, lengths within it, the songs past their end dropped
    for shape, lens, _ in model.calls:
        assert shape[1] == max(lens) <= 4 and len(lens) == shape[0] and min(lens) > 0
    assert abs(hist[0]['valid_log_likelihood'] - 5.0) < 1e-12


# ----------------------------------------------------------------------------- sample post-processing (row f4)
def _parse_smf(data):
    """Minimal Standard MIDI File reader for the round-trip test: returns (format, division, [track event lists]) with
    events (tick, status, data bytes)."""
    assert data[:4] == b'MThd' and int.from_bytes(data[4:8], 'big') == 6
    fmt, ntrk, div = (int.from_bytes(data[8 + 2 * i:10 + 2 * i], 'big') for i in range(3))
    pos, tracks = 14, []
    for _ in range(ntrk):
        assert data[pos:pos + 4] == b'MTrk'
        n = int.from_bytes(data[pos + 4:pos + 8], 'big')
        body, pos = data[pos + 8:pos + 8 + n], pos + 8 + n
        i, tick, ev = 0, 0, []
        while i < len(body):
            delta = 0
            while True:
                delta = (delta << 7) | (body[i] & 0x7F)
                i += 1
                if not body[i - 1] & 0x80:
                    break
            tick += delta
            st = body[i]
            if st == 0xFF:
                ln = body[i + 2]
                ev.append((tick, 0xFF, bytes(body[i + 1:i + 3 + ln])))
                i += 3 + ln
            elif st & 0xF0 == 0xC0:
                ev.append((tick, st, bytes(body[i + 1:i + 2])))
                i += 2
            else:
                ev.append((tick, st, bytes(body[i + 1:i + 3])))
                i += 3
        tracks.append(ev)
    assert pos == len(data)
    return fmt, div, tracks


def test_midi_export_round_trip(tmp_path):
    from multinn_b200.multinn import default_config
    from multinn_b200.utils import data as D
    cfg = default_config()['data']
    rng = np.random.default_rng(0)
    x = (rng.random((2, 40, 84, 5)) < 0.06).astype(np.float32)
    x[0, 38:40, 10, 1] = 1.0                                        # a note still sounding at the end of the song
    music = D.pad_to_midi(x, cfg)
    assert music.shape == (2, 40, 128, 5) and music[:, :, :24].sum() == 0 and music[:, :, 108:].sum() == 0
    paths = D.save_music(music, 1, cfg, 'unit', save_dir=str(tmp_path), song_labels=['t3'])
    assert [p.split('/')[-1] for p in paths] == ['unit_t3_0.mid', 'unit_t3_1.mid']          # data.py:196-200 names
    for s, path in enumerate(paths):
        fmt, div, tracks = _parse_smf(open(path, 'rb').read())
        assert fmt == 1 and div == cfg['beat_resolution'] and len(tracks) == 6
        assert tracks[0][0] == (0, 0xFF, b'\x51\x03' + (500000).to_bytes(3, 'big'))         # 120 bpm
        for m, name in enumerate(cfg['instruments']):
            ev = tracks[m + 1]
            prog = [e for e in ev if e[1] & 0xF0 == 0xC0]
            assert len(prog) == 1 and prog[0][2][0] == cfg['programs'][m]
            ch = prog[0][1] & 0x0F
            assert (ch == 9) == cfg['is_drums'][m]
            roll = np.zeros((40, 128), bool)
            sounding = {}
            gain = D.TRACK_GAIN.get(name, 1.0)
            for tick, st, d in ev:
                if st & 0xF0 == 0x90:
                    assert d[1] == int(round(100 * gain)) and d[0] not in sounding
                    sounding[d[0]] = tick
                elif st & 0xF0 == 0x80:
                    roll[sounding.pop(d[0]):tick, d[0]] = True
            assert not sounding
            np.testing.assert_array_equal(roll, music[s, :, :, m] > 0)


def test_evaluator_scores_bar_music():
    """model.evaluator() (multinn_core.py:343-362) = reshape to bars + metric summary; needs no device."""
    from multinn_b200.metrics import musical
    from multinn_b200.modes.core import MultINNCore
    from multinn_b200.multinn import default_config

    class Stub:
        _config = default_config()
        tracks = list(_config['data']['instruments'])
    x = (np.random.default_rng(1).random((2, 96, 84, 5)) < 0.05).astype(np.float32)
    got = MultINNCore.evaluator(Stub())(x)
    want = musical.metric_summary(x.reshape(2, 2, 48, 84, 5), Stub.tracks)
    assert got == want and len(got) == 29
    import torch
    assert MultINNCore.evaluator(Stub())(torch.from_numpy(x)) == want


def test_sample_songs_driver_with_stub_sampler(tmp_path):
    """sample.py:40-115 after the checkpoint load, with a sampler stub on the host (device='cpu'): tiling of the intros,
    intro + samples concatenation, 128-pitch padding, the saved ids / file names, and the metric table."""
    import torch
    from multinn_b200.multinn import default_config
    cfg = default_config()
    cfg['sampling'] = {'num_songs': 2, 'intro_beats': 2, 'sample_beats': 10, 'num_save': 2,
                       'intro_ids': {'train': {'start': 0, 'end': 2}, 'valid': {'start': 1, 'end': 2}},
                       'save_ids': {'train': [1], 'valid': [0]}}
    rng = np.random.default_rng(4)
    Xt = (rng.random((3, 30, 84, 5)) < 0.05).astype(np.float32)
    Xv = (rng.random((2, 30, 84, 5)) < 0.05).astype(np.float32)
    seen = {}

    class Model:
        def sampler(self, num_beats):
            steps = num_beats * 12

            def sample(x, u=None, seed=0):
                seen['intro'] = tuple(x.shape)
                g = torch.Generator().manual_seed(7)
                return (torch.rand(x.shape[0], steps, 84, 5, generator=g) < 0.04).float()
            return sample

    out = U.sample_songs(Model(), Xt, Xv, cfg, epoch=3, samples_dir=str(tmp_path), eval_samples=True, device='cpu',
                         name='unit')
    assert seen['intro'] == (6, 24, 84, 5)                          # (2 train + 1 valid intros) x num_songs, 2 beats
    assert out['samples'].shape == (6, 24 + 120, 128, 5)             # 3 bars of 48 steps
    np.testing.assert_array_equal(out['samples'][:3, :24, 24:108], np.concatenate([Xt[:2, :24], Xv[1:2, :24]]))
    assert [p.split('/')[-1] for p in out['paths']] == ['eval_unit_e3_t1_0.mid', 'eval_unit_e3_t1_1.mid',
                                                          'eval_unit_e3_v0_0.mid', 'eval_unit_e3_v0_1.mid']
    assert set(out['metrics']) == {'EB', 'UP', 'UPC', 'QN', 'PR', 'DP', 'TD'} and out['metrics']['TD'].shape == (4, 4)
    assert out['table'].count('\n') >= 10 and 'Drums' in out['table']


def test_recurrence_schedule_dispatch_for_the_baseline_shards():
    """Which schedule common/rnn.py picks for the per-GPU batches of the C5 scaling sweep (T = 256): full time-chunk
    pipeline at 256 / 512 (8 and 4 GPUs), forward-only pipeline at 1024 (2 GPUs), tail pipeline at 2048 (1 GPU), plain
    phase-by-phase for short or ragged time axes. Pure host logic: runs on a CPU arena."""
    from multinn_b200.multinn import MultINN, default_config, default_params
    model = MultINN(default_config(), default_params(mode='composer', num_hidden=128, num_hidden_rnn=(16, 8)), 'composer',
                    device='cpu')
    rnn = model.generators[0].rnn
    T = 256
    for B in (256, 512):
        assert rnn.use_pipeline(T, B) and rnn.use_any_pipeline(T, B) and rnn.use_fwd_pipeline(T, B)
    assert not rnn.use_pipeline(T, 1024) and rnn.use_fwd_pipeline(T, 1024)
    assert not rnn.use_pipeline(T, 2048) and not rnn.use_fwd_pipeline(T, 2048) and rnn.use_any_pipeline(T, 2048)
    for t, b in ((48, 256), (40, 256), (250, 2048)):            # T < 2 chunks, or not a multiple of the 32-step chunk
        assert not rnn.use_any_pipeline(t, b), (t, b)


def test_bf16_twin_lookup_matches_row_blocks_only():
    """ops._twin_of: a GEMM operand is redirected to the registered bf16 plane only when it is a full-width row block of the
    registered fp32 matrix (pointer arithmetic on views; no launch involved)."""
    import torch
    from multinn_b200 import ops
    base = torch.zeros(12, 20)
    twin = torch.zeros(12, 24, dtype=torch.int16)
    other = torch.zeros(12, 20)
    saved = list(ops._twins)
    try:
        del ops._twins[:]
        assert ops._twin_of(base) is None
        ops.register_twin(base, twin)
        assert ops._twin_of(base) == (twin.data_ptr(), 24)
        assert ops._twin_of(base[4:9]) == (twin.data_ptr() + 4 * 24 * 2, 24)
        assert ops._twin_of(base[:, :12]) is None          # narrower rows
        assert ops._twin_of(base[::2]) is None             # strided rows
        assert ops._twin_of(other[4:9]) is None            # another matrix
        assert ops._twin_of(base.view(24, 10)) is None     # different row structure
        ops.register_twin(base, torch.zeros(12, 32, dtype=torch.int16))   # re-registration replaces
        assert len(ops._twins) == 1 and ops._twin_of(base)[1] == 32
    finally:
        ops._twins[:] = saved


def test_bookkeeping_classes_behave_like_the_reference_ones():
    """tests/golden/ref_host.json was recorded from the reference's own TrainingStats / LossAccumulator
    (tools/make_golden_host.py); the rewritten classes must show the same counters, printed line and pickled tuple."""
    import json
    import os
    import pickle
    import tempfile
    from multinn_b200.utils.training import LossAccumulator, TrainingStats
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'ref_host.json')) as f:
        ref = json.load(f)
    acc = LossAccumulator()
    for x, want in zip(ref['losses'], ref['accumulator']):
        acc.update(float(x))
        assert repr(float(acc.loss())) == want['loss'] and acc.num_bad() == want['num_bad']
        assert acc.ratio_bad() == want['ratio_bad'] and str(acc) == want['line']
    acc.clear()
    assert repr(float(acc.loss())) == ref['cleared']['loss'] and acc.num_bad() == ref['cleared']['num_bad']
    st = TrainingStats()
    for op, want in zip(ref['stats_script'], ref['stats_states']):
        if isinstance(op, list):
            getattr(st, op[0])(*op[1:])
        else:
            getattr(st, op)()
        assert [st.steps, st.epoch, st.run, st.metric_best, st.idle_epochs] == want
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'stats.pkl')
        st.save(path)
        with open(path, 'rb') as f:
            assert list(pickle.load(f)) == ref['pickled']
        st2 = TrainingStats(steps=9, epoch=9, run=9, metric_best=0.0)
        st2.new_idle_epoch()
        st2.load(path)
        assert [st2.steps, st2.epoch, st2.run, st2.metric_best, st2.idle_epochs] == ref['loaded']


def test_data_helpers_behave_like_the_reference_ones(tmp_path):
    """tests/golden/ref_host_data.npz was recorded from the reference's own utils/data.py (tools/make_golden_host.py) on a
    small synthetic dataset: split, pixel padding + folding, lengths file, sampling inputs, MIDI padding."""
    import os
    import numpy as np
    from multinn_b200.utils import data as D
    ref = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_host_data.npz'))
    np.save(tmp_path / 'songs.npy', ref['songs'])
    np.save(tmp_path / 'lengths.npy', ref['lengths'])
    cfg = dict(filename=str(tmp_path / 'songs'), source='npy', sequence_lengths=None, instruments=['a', 'b', 'c', 'd', 'e'],
               split=dict(num_train=7, num_valid=3, num_test=2), pitch_range=dict(lowest=24, highest=30))
    for name, step, with_len in (('s1', 1, False), ('s3', 3, False), ('s2len', 2, True)):
        cfg['sequence_lengths'] = str(tmp_path / 'lengths.npy') if with_len else None
        (xt, lt), (xv, lv), (xs, ls) = D.load_data(cfg, step_size=step)
        for k, got in (('xt', xt), ('lt', lt), ('xv', xv), ('lv', lv), ('xs', xs), ('ls', ls)):
            np.testing.assert_array_equal(got, ref[f'{name}/{k}'], err_msg=f'{name}/{k}')
    cfg['sequence_lengths'] = None
    (xt, _), (xv, _), _ = D.load_data(cfg, step_size=1)
    samp = dict(intro_beats=2, intro_ids=dict(train=dict(start=1, end=5), valid=dict(start=0, end=2)),
                save_ids=dict(train=[0, 2], valid=[1]), num_save=3)
    intro, save_ids, labels = D.prepare_sampling_inputs(xt, xv, samp, beat_size=2)
    np.testing.assert_array_equal(intro, ref['samp/intro'])
    np.testing.assert_array_equal(save_ids, ref['samp/save_ids'])
    assert list(labels) == [str(x) for x in ref['samp/labels']]
    np.testing.assert_array_equal(D.pad_to_midi(xt[:2].astype(np.float32), cfg), ref['midi/pad'])
