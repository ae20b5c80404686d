"""CPU: multinn_b200/metrics/musical.py against golden vectors produced by the REFERENCE's own NumPy module
(tests/golden/musical_metrics.npz, tools/make_golden_musical.py) - the one row of the scope table whose reference code
runs in the build container, so this parity is pinned by the reference itself. Plus hand-computable known answers."""
import importlib.util
import os

import numpy as np
import pytest

from multinn_b200.metrics import musical as MM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('make_golden_musical', os.path.join(ROOT, 'tools', 'make_golden_musical.py'))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)
GOLD = np.load(os.path.join(ROOT, 'tests', 'golden', 'musical_metrics.npz'))


@pytest.mark.parametrize('case', gen.CASES, ids=[c[0] for c in gen.CASES])
def test_metrics_equal_reference_golden(case):
    name = case[0]
    roll = gen.case_roll(*case)
    assert int(roll.sum()) == int(GOLD[f'{name}/checksum'])
    chroma = MM.to_chroma(roll)
    np.testing.assert_array_equal(chroma, GOLD[f'{name}/chroma'])
    np.testing.assert_array_equal(MM.empty_bar_rate(roll), GOLD[f'{name}/EB'])
    np.testing.assert_array_equal(MM.num_pitches_used(roll), GOLD[f'{name}/UP'])
    np.testing.assert_array_equal(MM.num_pitches_used(chroma), GOLD[f'{name}/UPC'])
    np.testing.assert_array_equal(MM.qualified_note_rate(roll), GOLD[f'{name}/QN'])          # nan == nan here
    np.testing.assert_array_equal(MM.qualified_note_rate(roll, threshold=3), GOLD[f'{name}/QN3'])
    np.testing.assert_array_equal(MM.polyphonic_rate(roll), GOLD[f'{name}/PR'])
    np.testing.assert_allclose(MM.drum_in_pattern_rate(roll[..., 0]), float(GOLD[f'{name}/DP']), rtol=1e-12)
    np.testing.assert_allclose(MM.harmonicity(chroma), GOLD[f'{name}/TD'], rtol=1e-12, atol=1e-15)
    m = MM.sample_metrics(roll)
    np.testing.assert_array_equal(m['QN'], GOLD[f'{name}/QN_single'].astype(np.float32))
    assert MM.format_sample_metrics(m).split() == str(GOLD[f'{name}/table']).split()      # same numbers, same order
    # float {0,1} rolls (what the sampler returns) give the same numbers as bool rolls
    np.testing.assert_array_equal(MM.qualified_note_rate(roll.astype(np.float32)), GOLD[f'{name}/QN'])
    np.testing.assert_array_equal(MM.empty_bar_rate(roll.astype(np.float32)), GOLD[f'{name}/EB'])


def test_known_answers():
    roll = np.zeros((1, 2, 16, 24, 2), bool)
    roll[0, 0, 0:5, 3, 1] = True          # 5-step note
    roll[0, 0, 8:10, 3, 1] = True         # 2-step note (not > 2)
    roll[0, 0, 8:12, 4, 1] = True         # 4-step note
    roll[0, 0, 8:12, 5, 1] = True         # third simultaneous pitch on steps 8..9 -> polyphonic there
    roll[0, 0, 14:16, 7, 1] = True        # runs across the bar line into bar 1: one 4-step note
    roll[0, 1, 0:2, 7, 1] = True
    roll[0, 0, [0, 2, 4], 0, 0] = True    # drum hits on even steps of a 16-step bar: all on the (1, t) grid's ones
    np.testing.assert_array_equal(MM.empty_bar_rate(roll), [0.5, 0.0])
    np.testing.assert_array_equal(MM.num_pitches_used(roll), [0.5, 2.5])             # track 1: 4 pitches, then 1
    qn = MM.qualified_note_rate(roll)
    assert qn[1] == np.float32(4 / 5)     # 5 notes, 4 of them longer than 2 steps
    assert MM.polyphonic_rate(roll)[1] == (2 / 16) / 2
    assert MM.drum_in_pattern_rate(roll[..., 0]) == 1.0
    roll[0, 0, 1, 0, 0] = True            # an off-grid hit weighs the tolerance 0.1
    assert abs(MM.drum_in_pattern_rate(roll[..., 0]) - 3.1 / 4) < 1e-12
    assert MM.drum_in_pattern_rate(np.zeros((1, 1, 16, 5))) == 0.0
    # chroma quirk M1: 24 pitches -> classes of 2 consecutive pitches
    c = MM.to_chroma(roll)
    assert c.shape == (1, 2, 16, 12, 2) and c[0, 0, 8, 2, 1] == 2 and c[0, 0, 8, 1, 1] == 1   # pitches 4,5 | pitch 3
    # identical tracks are at tonal distance 0, the matrix is symmetric
    two = np.stack([roll[..., 1], roll[..., 1]], axis=-1)
    td = MM.harmonicity(MM.to_chroma(two))
    assert td.shape == (2, 2) and np.allclose(td, 0.0)


def test_argument_errors_match_reference():
    with pytest.raises(ValueError, match='5 dimensions'):
        MM.empty_bar_rate(np.zeros((2, 16, 84, 5)))
    with pytest.raises(ValueError, match='4 dimensions'):
        MM.drum_in_pattern_rate(np.zeros((1, 2, 16, 84, 5)))
    with pytest.raises(ValueError, match='chroma'):
        MM.harmonicity(np.zeros((1, 2, 16, 84, 5)))
    with pytest.raises(ValueError, match='Unsupported number of timesteps'):
        MM.drum_in_pattern_rate(np.zeros((1, 2, 20, 84)))


def test_metric_summary_layout():
    tracks = ['Drums', 'Piano', 'Guitar', 'Bass', 'Strings']
    roll = gen.case_roll(*gen.CASES[0])
    x = roll.reshape(roll.shape[0], -1, 84, 5).astype(np.float32)                   # model format [B, T, D, M]
    bars = MM.to_bars(x, beat_resolution=12, pitch_span=84)
    assert bars.shape == roll.shape
    s = MM.metric_summary(bars, tracks)
    assert len(s) == 2 * 5 + 1 + 3 * 4 + 6
    assert 'sample_scores/intra-track/Drums/DP' in s and 'sample_scores/intra-track/Drums/QN' not in s
    assert s['sample_scores/intra-track/Piano/QN'] == float(GOLD['sparse48/QN'][1])
    assert s['sample_scores/inter-track/TD/Piano-Bass/Piano-Bass'] == pytest.approx(float(GOLD['sparse48/TD'][1][3]), rel=1e-12)


@pytest.mark.skipif(not os.path.exists(gen.REF), reason='the reference checkout is not present on this machine')
def test_golden_file_is_what_the_reference_produces_now(tmp_path, monkeypatch):
    """Where /root/reference exists (the build container), re-run its musical.py and compare with the committed file."""
    monkeypatch.setattr(gen, 'ROOT', str(tmp_path))
    os.makedirs(tmp_path / 'tests' / 'golden')
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        gen.main()
    fresh = np.load(tmp_path / 'tests' / 'golden' / 'musical_metrics.npz')
    assert sorted(fresh.files) == sorted(GOLD.files)
    for k in GOLD.files:
        np.testing.assert_array_equal(fresh[k], GOLD[k], err_msg=k)


def test_metric_invariants_on_random_rolls():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.sampled_from([16, 24, 32, 48]), st.sampled_from([12, 20, 84]),
           st.floats(0.0, 0.6))
    def check(seed, steps, pitches, density):
        rng = np.random.default_rng(seed)
        roll = rng.random((2, 2, steps, pitches, 3)) < density
        chroma = MM.to_chroma(roll)
        assert chroma.shape[-2] == 12 and chroma.sum() == roll.sum()                 # folding keeps every note
        eb, up, pr = MM.empty_bar_rate(roll), MM.num_pitches_used(roll), MM.polyphonic_rate(roll)
        assert np.all((eb >= 0) & (eb <= 1)) and np.all((pr >= 0) & (pr <= 1)) and np.all((up >= 0) & (up <= pitches))
        assert np.all(MM.num_pitches_used(chroma) <= 12)
        dp = MM.drum_in_pattern_rate(roll[..., 0])
        assert 0.0 <= dp <= 1.0
        td = MM.harmonicity(chroma)
        finite = np.isfinite(td)
        assert np.allclose(td[finite], td.T[finite]) and np.all(np.diag(td)[np.isfinite(np.diag(td))] == 0)
        qn = MM.qualified_note_rate(roll, threshold=0)                               # every note lasts > 0 steps
        ok = np.isfinite(qn)
        assert np.all(qn[ok] >= 1.0)                                                 # == 1, or more with quirk M2
        silent = np.zeros_like(roll)
        assert np.all(MM.empty_bar_rate(silent) == 1) and np.all(np.isnan(MM.qualified_note_rate(silent)))

    check()


@pytest.mark.parametrize('name', ['sparse48', 'dense96', 'pitch20', 'silent_track'])
def test_metric_summary_equals_the_reference_tf_module(name):
    """tests/golden/musical_summary.json: tags and values recorded from the reference's `get_metric_summary_ops`
    (metrics/musical_tf.py, run on the NumPy TF stand-in by tools/make_golden_musical.py::summary_cases)."""
    import json
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'musical_summary.json')) as f:
        ref = json.load(f)[name]
    case = [c for c in gen.CASES if c[0] == name][0]
    roll = gen.case_roll(*case)
    with np.errstate(all='ignore'):
        got = MM.metric_summary(roll.astype(np.float32), ['Drums', 'Piano', 'Guitar', 'Bass', 'Strings'])
    assert sorted(got) == sorted(ref)
    for tag, want in ref.items():
        if want is None:
            assert np.isnan(got[tag]), tag
        else:
            assert got[tag] == pytest.approx(want, rel=1e-6, abs=1e-9), tag      # the TF module computes EB / UP / PR in fp32
