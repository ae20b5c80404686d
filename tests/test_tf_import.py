"""TF1-variable import / export (scope-table row f2), CPU: round trips through TF-shaped variables with shuffled order and
foreign scope prefixes, explicit name maps, and readable failures. (Unverified against a real TF checkpoint -- see the
module docstring of multinn_b200/utils/tf_import.py.)"""
import numpy as np
import pytest

from multinn_b200.multinn import MultINN, default_config, default_params
from multinn_b200.utils import tf_import as T


def _composer(seed):
    return MultINN(default_config(), default_params(mode='composer', num_hidden=32, num_hidden_rnn=(16, 8), keep_prob=1.0),
                   'composer', device='cpu', seed=seed)


def _same(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    return all(np.array_equal(sa[k].numpy(), sb[k].numpy()) for k in sa)


def test_composer_round_trip_with_foreign_prefixes_and_shuffled_order():
    src, dst = _composer(1), _composer(2)
    assert not _same(src.arena, dst.arena)
    tfv = T.export_tf_variables(src)
    assert tfv['generator/rnn/multi_rnn_cell/cell_1/cudnn_compatible_lstm_cell/kernel'].shape == (24, 32)
    assert tfv['generator/Piano/nade/w_enc'].shape == (84, 1, 32) and tfv['generator/Piano/nade/w_dec'].shape == (84, 32, 1)
    # whatever scope prefix TF really generated, only leaves + indices + track names matter
    renamed = {('MultINN/generators/rnn-multinade/all/' + n.split('/', 1)[1] + ':0'): a for n, a in tfv.items()}
    keys = list(renamed)
    np.random.default_rng(0).shuffle(keys)
    applied = T.load_tf_variables(dst, {k: renamed[k] for k in keys})
    assert _same(src.arena, dst.arena)
    assert len(applied['generator/nade/w_enc']) == 5 and 'Drums' in applied['generator/nade/w_enc'][0]


def test_jamming_tracks_are_matched_by_name():
    mk = lambda s: MultINN(default_config(), default_params(mode='jamming', num_hidden=32, num_hidden_rnn=(8,), keep_prob=1.0),
                           'jamming', device='cpu', seed=s)
    src, dst = mk(3), mk(4)
    tfv = T.export_tf_variables(src)
    keys = sorted(tfv, reverse=True)
    T.load_tf_variables(dst, {k: tfv[k] for k in keys})
    assert _same(src.arena, dst.arena)


def test_dbn_encoder_and_rnn_rbm_round_trip():
    mk = lambda s: MultINN(default_config(), default_params(mode='joint', encoder='DBN', encoder_hidden=[24, 12], generator='RBM',
                                                            num_hidden=16, num_hidden_rnn=(8,)), 'joint', device='cpu', seed=s)
    src, dst = mk(5), mk(6)
    T.load_tf_variables(dst, T.export_tf_variables(src))
    T.load_tf_variables(dst, T.export_tf_variables(src, 'encoders'), which='encoders')
    assert _same(src.arena, dst.arena) and _same(src.encoder_arena, dst.encoder_arena)
    assert 'encoder/all/rbm/1/W' in T.export_tf_variables(src, 'encoders')


def test_explicit_name_map_and_failures():
    src, dst = _composer(7), _composer(8)
    tfv = T.export_tf_variables(src)
    odd = dict(tfv)
    odd['some/other/name'] = odd.pop('generator/dense/kernel')
    with pytest.raises(ValueError, match='dense_kernel'):
        T.load_tf_variables(dst, odd)
    T.load_tf_variables(dst, odd, name_map={'generator/dense/kernel': 'some/other/name'})
    assert _same(src.arena, dst.arena)
    bad = dict(tfv)
    bad['generator/dense/bias'] = np.zeros(7, np.float32)
    with pytest.raises(ValueError, match='shape'):
        T.load_tf_variables(_composer(9), bad)
    extra = dict(tfv)
    extra['x/dense_3/kernel'] = np.zeros((2, 2), np.float32)
    with pytest.raises(ValueError, match='unused TF variables'):
        T.load_tf_variables(_composer(9), extra)
    T.load_tf_variables(_composer(9), extra, strict=False)


# ----------------------------------------------------------------------------- checkpoint files without TensorFlow
def test_checkpoint_file_round_trip_and_load(tmp_path):
    """utils/tf_checkpoint.py: TF V2 checkpoint (tensor bundle) files written and read in pure Python: many small index
    blocks (prefix-compressed keys, restart points), dtypes, scalars, the directory's `checkpoint` state file, and the
    whole path file -> variables -> model parameters."""
    from multinn_b200.utils import tf_checkpoint as C
    src, dst = _composer(11), _composer(12)
    tfv = {f'MultINN/generators/{n}': a for n, a in T.export_tf_variables(src).items()}
    tfv['global_step'] = np.array(1234, np.int64)                               # scalar, ignored by the import
    tfv['MultINN/generators/generator/dense/kernel/Adam'] = np.zeros((8, 1700), np.float32)    # optimiser slot, ignored
    tfv['flags'] = np.array([True, False, True])
    tfv['half'] = np.arange(6, dtype=np.float16).reshape(2, 3)
    prefix = str(tmp_path / 'generators' / 'model-77')
    C.write_checkpoint(prefix, tfv, block_size=200)                              # forces many data blocks
    assert sorted(f.name for f in (tmp_path / 'generators').iterdir()) == ['checkpoint', 'model-77.data-00000-of-00001',
                                                                           'model-77.index']
    for where in (prefix, prefix + '.index', str(tmp_path / 'generators')):
        got = C.read_checkpoint(where, verify_data=True)
        assert sorted(got) == sorted(tfv)
        for n in tfv:
            assert got[n].dtype == tfv[n].dtype and got[n].shape == tfv[n].shape and np.array_equal(got[n], tfv[n]), n
    only = C.read_checkpoint(prefix, names=lambda n: n.endswith('w_enc'))
    assert len(only) == 5
    T.load_tf_variables(dst, C.read_checkpoint(prefix), strict=False)
    assert _same(src.arena, dst.arena)


def test_checkpoint_reader_rejects_damage(tmp_path):
    from multinn_b200.utils import tf_checkpoint as C
    prefix = str(tmp_path / 'm')
    C.write_checkpoint(prefix, {'a/b': np.arange(12, dtype=np.float32).reshape(3, 4), 'a/c': np.ones(5, np.int32)})
    idx = bytearray(open(prefix + '.index', 'rb').read())
    bad = bytearray(idx)
    bad[3] ^= 0x40
    open(prefix + '.index', 'wb').write(bad)
    with pytest.raises(C.CheckpointFormatError, match='checksum|corrupt|truncated|varint'):
        C.read_checkpoint(prefix)
    bad = bytearray(idx)
    bad[-1] ^= 0xFF
    open(prefix + '.index', 'wb').write(bad)
    with pytest.raises(C.CheckpointFormatError, match='magic'):
        C.read_checkpoint(prefix)
    open(prefix + '.index', 'wb').write(idx)
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[5] ^= 1
    open(prefix + '.data-00000-of-00001', 'wb').write(data)
    assert C.read_checkpoint(prefix)['a/c'].sum() == 5                           # not verified by default
    with pytest.raises(C.CheckpointFormatError, match='tensor checksum'):
        C.read_checkpoint(prefix, verify_data=True)
    with pytest.raises(FileNotFoundError):
        C.read_checkpoint(str(tmp_path / 'missing'))


def test_crc32c_known_answers():
    from multinn_b200.utils.tf_checkpoint import crc32c, mask_crc
    assert crc32c(b'123456789') == 0xE3069283                                    # CRC-32C check value
    assert crc32c(bytes(32)) == 0x8a9136aa and crc32c(bytes([0xFF] * 32)) == 0x62a8ab43      # RFC 3720 B.4 vectors
    assert crc32c(b'world', crc32c(b'hello ')) == crc32c(b'hello world')
    assert mask_crc(crc32c(b'foo')) != crc32c(b'foo')


def test_model_save_tf_load_tf(tmp_path):
    mk = lambda s: MultINN(default_config(), default_params(mode='joint', encoder='DBN', encoder_hidden=[24, 12], generator='RBM',
                                                            num_hidden=16, num_hidden_rnn=(8,)), 'joint', device='cpu', seed=s)
    src, dst = mk(21), mk(22)
    src.save_tf(str(tmp_path / 'generators' / 'model'))
    src.save_tf(str(tmp_path / 'encoders' / 'model'), which='encoders')
    applied = dst.load_tf(str(tmp_path / 'generators'))
    dst.load_tf(str(tmp_path / 'encoders'), which='encoders', strict=True)
    assert _same(src.arena, dst.arena) and _same(src.encoder_arena, dst.encoder_arena)
    assert applied['generator/rbm/W'] == 'generator/rbm/W'


def test_checkpoint_table_round_trip_property(tmp_path_factory):
    """Random variable names with long shared prefixes, random block sizes (1 entry per block .. everything in one block),
    restart points every 16 keys: the table writer / reader pair keeps every key, value and order."""
    from hypothesis import given, settings, strategies as st
    from multinn_b200.utils import tf_checkpoint as C
    root = tmp_path_factory.mktemp('tables')
    counter = [0]

    @settings(max_examples=30, deadline=None)
    @given(st.sets(st.text(alphabet='abc/_0', min_size=1, max_size=24), min_size=1, max_size=60),
           st.integers(1, 3000), st.integers(0, 2 ** 31 - 1))
    def check(names, block_size, seed):
        rng = np.random.default_rng(seed)
        items = [(n.encode(), rng.bytes(int(rng.integers(0, 40)))) for n in names]
        path = str(root / f't{counter[0]}')
        counter[0] += 1
        C.write_table(path, items, block_size=block_size)
        assert C.read_table(path) == sorted(items)

    check()
