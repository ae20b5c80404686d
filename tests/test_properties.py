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
