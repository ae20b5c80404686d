"""TF1 checkpoint variables <-> multinn_b200 parameters (scope-table row f2; reference common/model.py:180-234 saves the
trainable variables of a model with tf.train.Saver, core/multinn_core.py:425-448 restores the encoders).

NOT VERIFIED AGAINST A REAL CHECKPOINT: tensorflow 1.13.1 cannot be installed here and the reference ships no
checkpoint, so the exact variable-scope prefixes TF generates from the reference's name_scope / variable_scope calls
(common/rnn.py:113, generators/rnn_estimator.py:86, common/nade.py:52, common/rbm.py:38, common/dbn.py:44-54) cannot be
confirmed. The import therefore does not depend on prefixes: TF variables are classified by their LEAF names and shapes
(which come from TF library code and from tf.Variable(name=...) literals in the reference), ordered by the layer / track
indices that appear in their names, and matched to this package's parameters of the same kind in the same order;
any ambiguity or shape mismatch raises with a readable report, and an explicit `name_map` overrides the inference.

The tensor layouts need no permutation: CudnnCompatibleLSTMCell's kernel is [(input + units), 4*units] with gate
blocks i, c~, f, o and rows [x ; h] -- exactly the layout the kernels here use; Dense kernels are
[in, units]; NADE w_enc [D,1,H] / w_dec [D,H,1] lose their singleton axis and are stacked over tracks; RBM W [D,H],
bh [1,H], bv [1,D] are identical.

Getting the variables out of a checkpoint: `utils/tf_checkpoint.read_checkpoint(ckpt_dir)` (pure Python), or in an
environment that has TensorFlow:
    r = tf.train.load_checkpoint(ckpt_dir); np.savez('vars.npz', **{n: r.get_tensor(n) for n in r.get_variable_to_shape_map()})
"""
import re

import numpy as np

# kind -> regex on the TF variable name (':0' suffixes are stripped first)
LEAF_PATTERNS = [
    ('lstm_kernel', re.compile(r'(?:^|/)cell_(\d+)/[a-z_]*lstm_cell/kernel$')),
    ('lstm_bias', re.compile(r'(?:^|/)cell_(\d+)/[a-z_]*lstm_cell/bias$')),
    ('dense_kernel', re.compile(r'(?:^|/)dense(?:_(\d+))?/kernel$')),
    ('dense_bias', re.compile(r'(?:^|/)dense(?:_(\d+))?/bias$')),
    ('w_enc', re.compile(r'(?:^|/)w_enc$')),
    ('w_dec', re.compile(r'(?:^|/)w_dec$')),
    ('rbm_W', re.compile(r'(?:^|/)rbm(?:/(\d+))?/W$')),
    ('rbm_bh', re.compile(r'(?:^|/)rbm(?:/(\d+))?/bh$')),
    ('rbm_bv', re.compile(r'(?:^|/)rbm(?:/(\d+))?/bv$')),
    ('Wuh', re.compile(r'(?:^|/)Wuh$')),
    ('Wuv', re.compile(r'(?:^|/)Wuv$')),
]
# kind of each parameter of this package, by the leaf of its arena name
OUR_PATTERNS = [
    ('lstm_kernel', re.compile(r'/cell_(\d+)/kernel$')), ('lstm_bias', re.compile(r'/cell_(\d+)/bias$')),
    ('dense_kernel', re.compile(r'/dense(?:_(\d+))?/kernel$')), ('dense_bias', re.compile(r'/dense(?:_(\d+))?/bias$')),
    ('w_enc', re.compile(r'/nade/w_enc$')), ('w_dec', re.compile(r'/nade/w_dec$')),
    ('rbm_W', re.compile(r'/rbm(?:_(\d+))?/W$')), ('rbm_bh', re.compile(r'/rbm(?:_(\d+))?/bh$')),
    ('rbm_bv', re.compile(r'/rbm(?:_(\d+))?/bv$')), ('Wuh', re.compile(r'/Wuh$')), ('Wuv', re.compile(r'/Wuv$')),
]


def _classify(name, patterns):
    for kind, rx in patterns:
        m = rx.search(name)
        if m:
            idx = int(m.group(1)) if m.groups() and m.group(1) is not None else 0
            return kind, idx
    return None, 0


def _track_index(name, tracks):
    """Position of the first track name that appears as a path component of `name` (-1: none)."""
    parts = name.split('/')
    for i, t in enumerate(tracks):
        if t in parts:
            return i
    return -1


def _clean(name):
    return name[:-2] if name.endswith(':0') else name


def export_tf_variables(model, which='generators'):
    """{TF-style name: array in the TF shape} for the generator (or encoder) parameters: the inverse of
    `load_tf_variables`. The prefixes are this package's arena names (see the module docstring on prefixes)."""
    arena = model.arena if which == 'generators' else model.encoder_arena
    tracks = list(model.tracks)
    out = {}
    for name, t in arena.state_dict().items():
        a = t.numpy()
        kind, idx = _classify(name, OUR_PATTERNS)
        prefix = name.rsplit('/', 2)[0] if kind in ('lstm_kernel', 'lstm_bias') else name.rsplit('/', 1)[0]
        if kind in ('lstm_kernel', 'lstm_bias'):
            out[f'{prefix}/multi_rnn_cell/cell_{idx}/cudnn_compatible_lstm_cell/{"kernel" if kind == "lstm_kernel" else "bias"}'] = a
        elif kind in ('w_enc', 'w_dec'):
            base = prefix.rsplit('/', 1)[0]
            for m in range(a.shape[0]):
                scope = f'{base}/{tracks[m]}/nade' if a.shape[0] > 1 else f'{base}/nade'
                out[f'{scope}/{kind}'] = a[m][:, None, :] if kind == 'w_enc' else a[m][:, :, None]
        elif kind in ('rbm_W', 'rbm_bh', 'rbm_bv'):
            base = name.rsplit('/', 2)[0]
            leaf = kind.split('_')[1]
            has_idx = re.search(r'/rbm_(\d+)/', name) is not None
            out[f'{base}/rbm/{idx}/{leaf}' if has_idx else f'{base}/rbm/{leaf}'] = a
        else:
            out[name] = a
    return out


def load_tf_variables(model, variables, which='generators', name_map=None, strict=True):
    """Copies TF variables {name: ndarray} into the generator (or encoder) parameters of `model`.
    name_map: optional {arena parameter name: TF variable name} (for stacked NADE weights: a list of M TF names);
    entries given there are taken as is, the rest is inferred (module docstring). strict: every parameter must be
    found and every classified TF variable used. Returns {arena name: TF name(s)} as applied."""
    arena = model.arena if which == 'generators' else model.encoder_arena
    tracks = list(model.tracks)
    tf_vars = {_clean(n): np.asarray(a) for n, a in variables.items()}
    name_map = dict(name_map or {})
    pools = {}
    for n in sorted(tf_vars):
        kind, idx = _classify(n, LEAF_PATTERNS)
        if kind is not None:
            pools.setdefault(kind, []).append((_track_index(n, tracks), idx, n))
    for kind in pools:
        pools[kind].sort()
    used, applied, problems = set(), {}, []
    params = arena.named()

    def take(kind, want_shape, label):
        for j, (_, _, n) in enumerate(pools.get(kind, [])):
            if n in used:
                continue
            if tuple(np.squeeze(tf_vars[n]).shape) == tuple(np.squeeze(np.empty(want_shape)).shape):
                used.add(n)
                return n
        problems.append(f'{label}: no unused TF variable of kind {kind} with shape {tuple(want_shape)} '
                        f'(candidates: {[(n, tf_vars[n].shape) for _, _, n in pools.get(kind, []) if n not in used]})')
        return None

    # parameters in arena order = generator / track order, then layer order: the same order the pools are sorted in
    for pname, p in params.items():
        shape = tuple(p.data.shape)
        kind, _ = _classify(pname, OUR_PATTERNS)
        if pname in name_map:
            names = name_map[pname]
            names = [names] if isinstance(names, str) else list(names)
            arrs = [np.squeeze(tf_vars[_clean(n)]) for n in names]
            used.update(_clean(n) for n in names)
            arena.load(pname, np.stack(arrs) if len(arrs) > 1 or (kind in ('w_enc', 'w_dec')) else arrs[0])
            applied[pname] = names
            continue
        if kind is None:
            problems.append(f'{pname}: no TF leaf pattern for this parameter; pass it in name_map')
            continue
        if kind in ('w_enc', 'w_dec'):                       # [M, D, H] <- M variables [D,1,H] / [D,H,1]
            got = [take(kind, shape[1:], f'{pname}[{m}]') for m in range(shape[0])]
            if all(g is not None for g in got):
                arena.load(pname, np.stack([np.squeeze(tf_vars[g]).reshape(shape[1:]) for g in got]))
                applied[pname] = got
        else:
            g = take(kind, shape, pname)
            if g is not None:
                arena.load(pname, tf_vars[g].reshape(shape))
                applied[pname] = g
    if strict:
        left = [n for k in pools for _, _, n in pools[k] if n not in used]
        if left:
            problems.append(f'unused TF variables: {left}')
    if problems:
        raise ValueError('TF variable import failed:\n  ' + '\n  '.join(problems))
    return applied
