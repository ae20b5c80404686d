"""Generates tests/golden/ref_host.json by running the REFERENCE'S OWN host bookkeeping classes
(/root/reference/multinn/utils/training.py: TrainingStats, LossAccumulator -- imported unmodified; the module's `import
tensorflow` resolves to tests/tf_stub) on a scripted sequence of calls; tests/test_host_utils.py replays the script on
multinn_b200.utils.training and compares every observable (counters, printed line, pickled tuple).
  python tools/make_golden_host.py"""
import json
import os
import pickle
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'tf_stub'))
sys.path.insert(0, '/root/reference/multinn')

from utils.training import LossAccumulator, TrainingStats  # noqa: E402
from utils import data as ref_data  # noqa: E402

import numpy as np  # noqa: E402

LOSSES = [61.5, 58.25, float('nan'), 40.125, float('inf'), float('-inf'), 33.0, float('nan'), 12.5]
STATS_SCRIPT = ['new_step', 'new_step', 'new_epoch', ('update_metric_best', 41.5), 'new_idle_epoch', 'new_idle_epoch',
                'new_run', 'reset_idle_epochs', 'new_step', 'new_idle_epoch']


def main():
    acc = LossAccumulator()
    trace = []
    for x in LOSSES:
        acc.update(x)
        trace.append(dict(loss=repr(float(acc.loss())), num_bad=acc.num_bad(), ratio_bad=acc.ratio_bad(), line=str(acc)))
    acc.clear()
    cleared = dict(loss=repr(float(acc.loss())), num_bad=acc.num_bad())
    st = TrainingStats()
    states = []
    for op in STATS_SCRIPT:
        if isinstance(op, tuple):
            getattr(st, op[0])(*op[1:])
        else:
            getattr(st, op)()
        states.append([st.steps, st.epoch, st.run, st.metric_best, st.idle_epochs])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'stats.pkl')
        st.save(path)
        with open(path, 'rb') as f:
            saved = list(pickle.load(f))
        st2 = TrainingStats(steps=9, epoch=9, run=9, metric_best=0.0)
        st2.new_idle_epoch()
        st2.load(path)
        loaded = [st2.steps, st2.epoch, st2.run, st2.metric_best, st2.idle_epochs]
    out = dict(losses=[repr(x) for x in LOSSES], accumulator=trace, cleared=cleared, stats_script=STATS_SCRIPT,
               stats_states=states, pickled=saved, loaded=loaded)
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_host.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1)
    print(path)
    data_cases()


def data_cases():
    """utils/data.py: load_data (split, pixel padding + reshape, lengths file), prepare_sampling_inputs, pad_to_midi."""
    rng = np.random.default_rng(3)
    res = {}
    with tempfile.TemporaryDirectory() as d:
        songs = (rng.random((14, 10, 6, 5)) < 0.3)                       # 10 time steps: not a multiple of step_size 3
        lengths = rng.integers(1, 4, size=14)
        np.save(os.path.join(d, 'songs.npy'), songs)
        np.save(os.path.join(d, 'lengths.npy'), lengths)
        cfg = dict(filename=os.path.join(d, 'songs'), source='npy', sequence_lengths=None,
                   instruments=['a', 'b', 'c', 'd', 'e'], split=dict(num_train=7, num_valid=3, num_test=2),
                   pitch_range=dict(lowest=24, highest=30))
        res['songs'], res['lengths'] = songs, lengths
        for name, step, with_len in (('s1', 1, False), ('s3', 3, False), ('s2len', 2, True)):
            cfg['sequence_lengths'] = os.path.join(d, 'lengths.npy') if with_len else None
            (xt, lt), (xv, lv), (xs, ls) = ref_data.load_data(cfg, step_size=step)
            res.update({f'{name}/xt': xt, f'{name}/lt': lt, f'{name}/xv': xv, f'{name}/lv': lv, f'{name}/xs': xs, f'{name}/ls': ls})
        cfg['sequence_lengths'] = None
        (xt, _), (xv, _), _ = ref_data.load_data(cfg, step_size=1)
        samp = dict(intro_beats=2, intro_ids=dict(train=dict(start=1, end=5), valid=dict(start=0, end=2)),
                    save_ids=dict(train=[0, 2], valid=[1]), num_save=3)
        intro, save_ids, labels = ref_data.prepare_sampling_inputs(xt, xv, samp, beat_size=2)
        res['samp/intro'], res['samp/save_ids'] = intro, save_ids
        res['samp/labels'] = np.array(labels)
        res['midi/pad'] = ref_data.pad_to_midi(xt[:2].astype(np.float32), cfg)
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_host_data.npz')
    np.savez_compressed(path, **res)
    print(path, len(res), 'arrays')


if __name__ == '__main__':
    main()
