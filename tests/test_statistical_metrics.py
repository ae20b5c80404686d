"""CPU: streaming statistical metrics (reference metrics/statistical.py:6-47) against brute force over the whole set, and
collect_metrics(classification=True) with a stub model."""
import numpy as np
import torch

from multinn_b200.metrics.statistical import BaseMetrics
from multinn_b200.utils import training as U


def test_streaming_equals_whole_set_and_quirk_f1():
    rng = np.random.default_rng(0)
    m = BaseMetrics()
    T_, P_, L_ = [], [], []
    for n in (5, 1, 9):
        t = rng.random((n, 7, 3)) < 0.3
        p = np.where(rng.random((n, 7, 3)) < 0.8, t, ~t)
        lp = rng.random((n, 3)) * 4
        m.update(torch.from_numpy(lp), torch.from_numpy(t.astype(np.float32)), torch.from_numpy(p))
        T_.append(t.reshape(-1)); P_.append(p.reshape(-1)); L_.append(lp.reshape(-1))
    t, p, lp = np.concatenate(T_), np.concatenate(P_), np.concatenate(L_)
    r = m.result()
    tp, fp, fn = (t & p).sum(), (~t & p).sum(), (t & ~p).sum()
    assert abs(r['log_likelihood'] - lp.mean()) < 1e-12 and abs(r['perplexity'] - np.exp(lp).mean()) < 1e-12
    assert abs(r['accuracy'] - (t == p).mean()) < 1e-12
    assert abs(r['precision'] - tp / (tp + fp)) < 1e-12 and abs(r['recall'] - tp / (tp + fn)) < 1e-12
    assert r['f1_score'] == r['precision']                       # quirk Q5 (statistical.py:37-38)
    pr, rc = tp / (tp + fp), tp / (tp + fn)
    assert abs(r['true_f1'] - 2 * pr * rc / (pr + rc)) < 1e-12 and r['rows'] == lp.size
    empty = BaseMetrics()
    empty.update(torch.zeros(4), torch.zeros(4, 2), torch.zeros(4, 2))
    assert empty.result()['precision'] == 0.0 and empty.result()['f1_score'] == 0.0 and empty.result()['accuracy'] == 1.0


class _Stub:
    """evaluate(): NLL = number of notes in the row, cond_probs = 0.9 where the target is on except on pitch 0."""

    def evaluate(self, x, lengths=None, cond_probs=False):
        B, T = x.shape[:2]
        keep = (torch.arange(T)[None, :] < lengths[:, None]).reshape(-1)
        rows = x.float().reshape(B * T, *x.shape[2:])[keep]
        out = {'nll': rows.sum(1)}
        if cond_probs:
            cp = rows * 0.9
            cp[:, 0] = 0.1
            out['cond_probs'] = cp
        return out


def test_collect_metrics_classification_with_variable_lengths():
    rng = np.random.default_rng(3)
    X = (rng.random((5, 8, 6, 2)) < 0.4).astype(np.uint8)
    lengths = np.array([8, 3, 8, 5, 8])
    got = U.collect_metrics(_Stub(), X, lengths, batch_size=2, piece_size=8, device='cpu', classification=True)
    rows = np.concatenate([X[b, :l] for b, l in enumerate(lengths)]).astype(bool)
    pred = rows.copy()
    pred[:, 0] = False
    tp, fp, fn = (rows & pred).sum(), (~rows & pred).sum(), (rows & ~pred).sum()
    assert got['rows'] == rows.shape[0] * 2
    assert abs(got['accuracy'] - (rows == pred).mean()) < 1e-12
    assert got['precision'] == 1.0 and abs(got['recall'] - tp / (tp + fn)) < 1e-12 and fp == 0
    plain = U.collect_metrics(_Stub(), X, lengths, batch_size=2, piece_size=8, device='cpu')
    assert plain['log_likelihood'] == got['log_likelihood'] and 'accuracy' not in plain


def test_streaming_result_does_not_depend_on_the_batching():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=30, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.lists(st.integers(1, 7), min_size=1, max_size=6))
    def check(seed, sizes):
        rng = np.random.default_rng(seed)
        n = sum(sizes)
        t = torch.from_numpy(rng.random((n, 5)) < 0.4)
        p = torch.from_numpy(rng.random((n, 5)) < 0.5)
        lp = torch.from_numpy(rng.random(n) * 3)
        whole = BaseMetrics()
        whole.update(lp, t, p)
        parts, a = BaseMetrics(), 0
        for s in sizes:
            parts.update(lp[a:a + s], t[a:a + s], p[a:a + s])
            a += s
        rw, rp = whole.result(), parts.result()
        assert rw.keys() == rp.keys()
        for k in rw:
            assert abs(rw[k] - rp[k]) < 1e-12, k

    check()
