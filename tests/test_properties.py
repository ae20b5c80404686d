"""Property tests (hypothesis) on the shape / length-mask logic of the host side: the padded flatten of
utils/sequences.py:6-37 and the batch pieces of the train / eval loops, for arbitrary shapes and lengths."""
import numpy as np
from hypothesis import given, settings, strategies as st

from multinn_b200.utils import training as U
from oracle import np_oracle as O


@st.composite
def lengths_case(draw):
    B = draw(st.integers(1, 6))
    T = draw(st.integers(1, 9))
    lengths = draw(st.lists(st.integers(1, T), min_size=B, max_size=B))
    if draw(st.booleans()):
        lengths[draw(st.integers(0, B - 1))] = T          # max(lengths) == T, as dynamic_decode assumes
    return B, T, np.array(lengths)


@settings(max_examples=60, deadline=None)
@given(lengths_case())
def test_flatten_valid_rows_is_the_sequence_mask_in_batch_major_order(case):
    B, T, lengths = case
    rows = O.flatten_valid_rows(lengths, T)
    mask = np.arange(T)[None, :] < lengths[:, None]                 # tf.sequence_mask
    np.testing.assert_array_equal(rows, np.flatnonzero(mask.reshape(-1)))   # tf.where order = row-major = b*T + t
    assert len(rows) == lengths.sum() and np.all(np.diff(rows) > 0)
    # time-major device rows t*B + b hold the same set
    tm = np.flatnonzero((np.arange(T)[:, None] < lengths[None, :]).reshape(-1))
    assert sorted((r % T) * B + r // T for r in rows) == sorted(tm)


@settings(max_examples=60, deadline=None)
@given(lengths_case(), st.integers(1, 4), st.integers(1, 5), st.integers(0, 2 ** 31 - 1))
def test_training_pieces_cover_every_valid_frame_exactly_once(case, batch_size, piece_size, seed):
    B, T, lengths = case
    X = np.arange(B * T, dtype=np.int64).reshape(B, T, 1, 1)        # every frame carries its own id
    ids = np.random.default_rng(seed).permutation(B)
    seen = []
    for bi, songs, lens in U.training_pieces(X, lengths, ids, batch_size, piece_size):
        assert songs.shape[0] == len(lens) and songs.shape[1] == lens.max() <= piece_size and lens.min() >= 1
        for row, n in zip(songs, lens):
            seen.extend(row[:n, 0, 0].tolist())
    expect = [b * T + t for b in range(B) for t in range(lengths[b])]
    assert sorted(seen) == expect


@settings(max_examples=40, deadline=None)
@given(lengths_case(), st.integers(1, 4), st.integers(1, 5))
def test_evaluation_pieces_shapes(case, batch_size, piece_size):
    B, T, lengths = case
    X = np.zeros((B, T, 2, 1), dtype=np.float32)
    n_batches = -(-B // batch_size)
    n_pieces = -(-T // piece_size)
    pieces = list(U.evaluation_pieces(X, lengths, batch_size, piece_size))
    assert len(pieces) == n_batches * n_pieces
    for songs, seq in pieces:
        assert songs.shape[0] == len(seq) and 1 <= seq.max() <= songs.shape[1] <= piece_size


def _nade_segment_form(x, b_enc, b_dec, w_enc, w_dec):
    """The algorithm of the CUDA kernels restated in NumPy: the hidden state of a row only changes after a set target
    bit, so sigmoid(a) is re-evaluated only for the rows that start a new segment at dim i (nade.cu header). Same array
    shapes in the dot products as the loop form, so that BLAS sums in the same order."""
    N, D = x.shape
    log_p = np.zeros(N)
    p = np.zeros((N, D))
    a = b_enc.copy()
    h = O.sigmoid(a)
    evals = N
    for i in range(D):
        if i > 0:
            rows = np.flatnonzero(x[:, i - 1] == 1.0)          # a new segment starts below a set bit
            if len(rows):
                a[rows] = a[rows] + w_enc[i - 1][None, :]
                h[rows] = O.sigmoid(a[rows])
                evals += len(rows)
        l = b_dec[:, i] + h @ w_dec[i]
        p[:, i] = O.sigmoid(l)
        log_p = log_p + (x[:, i] * O.safe_log(p[:, i]) + (1 - x[:, i]) * O.safe_log(1 - p[:, i]))
    return -log_p, p, evals


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 6), st.integers(1, 12), st.integers(1, 6), st.floats(0.0, 1.0), st.integers(0, 2 ** 31 - 1))
def test_segment_form_equals_the_reference_loop_bit_for_bit(N, D, H, density, seed):
    """common/nade.py:199-226 evaluates h = sigmoid(a_i) at every i; a_i only changes where v_{i-1} = 1, so evaluating it
    once per segment gives bit-identical results with (1 + popcount(v[:-1])) instead of D sigmoid vectors per row."""
    rng = np.random.default_rng(seed)
    x = (rng.random((N, D)) < density).astype(np.float64)
    b_enc, b_dec = rng.standard_normal((N, H)), rng.standard_normal((N, D))
    w_enc, w_dec = rng.standard_normal((D, H)), rng.standard_normal((D, H))
    ref_nll, ref_p = O.nade_log_prob(x, b_enc, b_dec, w_enc, w_dec)
    nll, p, evals = _nade_segment_form(x, b_enc, b_dec, w_enc, w_dec)
    np.testing.assert_array_equal(p, ref_p)
    np.testing.assert_array_equal(nll, ref_nll)
    assert evals == N + int(x[:, :-1].sum())
    tri_nll, _ = O.nade_log_prob_triangular(x, b_enc, b_dec, w_enc, w_dec)
    np.testing.assert_allclose(tri_nll, ref_nll, rtol=1e-10, atol=1e-12)
