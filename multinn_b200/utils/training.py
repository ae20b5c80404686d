"""Training / evaluation host loop (mirrors reference multinn/utils/training.py:8-240 and the epoch loop of
multinn/train.py:153-282). The reference drives a TF session with feed_dict batches; here `model` is a
multinn_b200.MultINN and every batch piece is one call of the step / evaluate functions on device tensors.
Logging, TensorBoard summaries, MIDI export and the CLI are out of scope (control plane)."""
import pickle

import numpy as np


class TrainingStats:
    """Progress counters of a training run (role of reference utils/training.py:8-95): optimisation steps, epochs, runs
    (how often training was (re)started), the best monitored metric so far and the epochs since it last improved.
    On disk: the reference's pickle of the 4-tuple (steps, epoch, run, metric_best), so stats files interchange."""
    _PERSISTED = ('steps', 'epoch', 'run', 'metric_best')

    def __init__(self, steps=0, epoch=0, run=0, metric_best=1e3):
        self._c = dict(steps=steps, epoch=epoch, run=run, metric_best=metric_best, idle_epochs=0)

    def __getattr__(self, name):
        c = self.__dict__.get('_c', {})
        if name in c:
            return c[name]
        raise AttributeError(name)

    def _bump(self, key):
        self._c[key] += 1

    def new_step(self):
        self._bump('steps')

    def new_epoch(self):
        self._bump('epoch')

    def new_run(self):
        self._bump('run')

    def new_idle_epoch(self):
        self._bump('idle_epochs')

    def reset_idle_epochs(self):
        self._c['idle_epochs'] = 0

    def update_metric_best(self, val):
        self._c['metric_best'] = val

    def save(self, filename):
        with open(filename, 'wb') as fh:
            pickle.dump(tuple(self._c[k] for k in self._PERSISTED), fh)

    def load(self, filename):
        with open(filename, 'rb') as fh:
            self._c.update(zip(self._PERSISTED, pickle.load(fh)))


class LossAccumulator:
    """Running mean of the per-step training losses that leaves non-finite values out of the mean and tallies them by
    kind (role of reference utils/training.py:98-148; train.py:192 feeds it, train.py:199 prints it)."""
    _KINDS = ('nan', '+inf', '-inf')

    def __init__(self):
        self.clear()

    def clear(self):
        self._total, self._finite = 0.0, 0
        self._bad = dict.fromkeys(self._KINDS, 0)

    @staticmethod
    def _kind(x):
        if x != x:
            return 'nan'
        if x in (float('inf'), float('-inf')):
            return '+inf' if x > 0 else '-inf'
        return None

    def update(self, loss):
        kind = self._kind(float(loss))
        if kind is None:
            self._total += float(loss)
            self._finite += 1
        else:
            self._bad[kind] += 1

    def loss(self):
        return self._total / self._finite if self._finite else float('nan')

    def num_bad(self):
        return sum(self._bad.values())

    def ratio_bad(self):
        return self.num_bad() / (self._finite + self.num_bad())

    def __str__(self):
        tally = ', '.join(f'{k}: {self._bad[k]}' for k in self._KINDS)
        return f' - loss: {self.loss():7.3f} ({tally}, bad: {100. * self.ratio_bad():.2f}%)'


def training_pieces(X, lengths, ids, batch_size, piece_size):
    """The batch pieces one epoch feeds (train.py:165-178): for every batch of shuffled song ids and every piece offset
    j, the songs that still have frames past j, cut to the longest remaining length (at most `piece_size`).
    Yields (batch_index, songs[b, max_length, D, M], lengths[b]); `batch_index` changes once per batch (quirk Q11:
    the step counter advances per batch, not per piece)."""
    for bi, i in enumerate(range(0, X.shape[0], batch_size)):
        sel = ids[i:i + batch_size]
        for j in range(0, X.shape[1], piece_size):
            len_batch = lengths[sel] - j
            non_empty = np.where(len_batch > 0)[0]
            if len(non_empty) > 0:
                len_batch = np.minimum(len_batch[non_empty], piece_size)
                max_length = int(len_batch.max())
                yield bi, X[sel, j:j + max_length, ...][non_empty], len_batch


def evaluation_pieces(data, data_lengths, batch_size, piece_size):
    """The pieces collect_metrics feeds (training.py:201-213). Quirk Q9 kept: the piece offset j is NOT subtracted from
    the lengths, so every piece is evaluated over min(length, piece_size) frames. Lengths are capped at the frames the
    slice really has (a ragged last piece)."""
    for i in range(0, data.shape[0], batch_size):
        for j in range(0, data.shape[1], piece_size):
            seq = np.minimum(data_lengths[i:i + batch_size], piece_size)
            max_length = int(seq.max())
            songs = data[i:i + batch_size, j:j + max_length, :]
            yield songs, np.minimum(seq, songs.shape[1])


def _to_device(songs, device):
    import torch
    a = np.ascontiguousarray(songs)
    if a.dtype not in (np.uint8, np.bool_, np.float32):
        a = a.astype(np.float32)
    return torch.from_numpy(a).to(device, non_blocking=True)


def collect_metrics(model, data, data_lengths, batch_size, piece_size, device='cuda', distributed=None,
                    classification=False):
    """Streaming evaluation over a dataset (training.py:180-213 + metrics/statistical.py:22-34): the tf.metrics
    accumulators become counters on the device (metrics/statistical.py::BaseMetrics). Returns {'log_likelihood': mean
    NLL over all evaluated rows and tracks, 'perplexity': mean exp(NLL), 'rows': count}; with `classification=True`
    (models whose evaluate() takes `cond_probs=True`) also accuracy / precision / recall / f1_score of the thresholded
    conditional probabilities (predictions = p >= .5, rnn_multinade.py:139-143) against the targets.
    Data parallel (`distributed=None`: whenever torch.distributed is initialised with more than one rank): rank r
    evaluates the batches b with b % world == r and the accumulators are summed over the ranks with one allreduce, so
    every rank returns the metrics of the WHOLE dataset."""
    import torch
    import torch.distributed as dist
    from ..metrics.statistical import BaseMetrics
    rank, world = 0, 1
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if distributed:
        rank, world = dist.get_rank(), dist.get_world_size()
    acc = BaseMetrics(device)
    pieces_per_batch = len(range(0, data.shape[1], piece_size))
    for idx, (songs, seq) in enumerate(evaluation_pieces(data, data_lengths, batch_size, piece_size)):
        if (idx // pieces_per_batch) % world != rank:
            continue
        x = _to_device(songs, device)
        lengths = torch.as_tensor(np.asarray(seq))
        if not classification:
            acc.update(model.evaluate(x, lengths=lengths)['nll'])
            continue
        out = model.evaluate(x, lengths=lengths, cond_probs=True)
        B, T = x.shape[0], x.shape[1]
        targets = x.reshape(B * T, *x.shape[2:])                       # rows n = b*T + t, like the per-row results
        if out['cond_probs'].shape[0] != B * T:                        # padded rows were dropped (utils/sequences.py:29-37)
            keep = (torch.arange(T)[None, :] < lengths[:, None]).reshape(-1).nonzero().squeeze(1).to(x.device)
            targets = targets[keep]
        acc.update(out['nll'], targets, out['cond_probs'] >= 0.5)
    if distributed:
        acc.allreduce()
    res = acc.result()
    res.pop('loss')
    return res


def generate_music(model, sampler, intro_songs, num_songs=5, concat=True, device='cuda', u=None, seed=None):
    """training.py:216-240: tile the intros `num_songs` times, sample, optionally prepend the intro."""
    intro = np.tile(intro_songs, (num_songs, 1, 1, 1))
    samples = sampler(_to_device(intro, device), u=u, seed=seed).cpu().numpy()
    return np.concatenate([intro.astype(samples.dtype), samples], axis=1) if concat else samples


def sample_songs(model, X_train, X_valid, config, epoch=0, samples_dir=None, eval_samples=False, device='cuda', u=None,
                 seed=None, name='MultINN'):
    """The sampling run of sample.py:40-115 after the checkpoint is loaded: intros from the train/valid splits ->
    `model.sampler(sample_beats)` -> generate_music (intro + samples) -> pad_to_midi -> MIDI files of the `save_ids`
    (when `samples_dir` is given) -> musical metrics of all samples (when `eval_samples`).
    Returns dict(samples[N, steps, 128, tracks], paths, metrics, table)."""
    from ..metrics import musical
    from .data import pad_to_midi, prepare_sampling_inputs, save_music
    dc, sc = config['data'], config['sampling']
    beat_size = float(dc['beat_resolution'] / config['training']['num_pixels'])    # sample.py:39: steps per beat
    intro_songs, save_ids, song_labels = prepare_sampling_inputs(X_train, X_valid, sc, beat_size)
    sampler = model.sampler(num_beats=sc['sample_beats'])
    music = generate_music(model, sampler, intro_songs, num_songs=sc['num_songs'], device=device, u=u, seed=seed)
    samples = pad_to_midi(music, dc)
    out = {'samples': samples, 'paths': [], 'metrics': None, 'table': None}
    if samples_dir is not None:
        out['paths'] = save_music(samples[save_ids], num_intro=len(save_ids) // sc['num_save'], data_config=dc,
                                  base_path=f'eval_{name}_e{epoch}', save_dir=samples_dir, song_labels=song_labels)
    if eval_samples:
        bars = samples.reshape((samples.shape[0], -1, dc['beat_resolution'] * 4) + samples.shape[-2:])
        out['metrics'] = musical.sample_metrics(bars)
        out['table'] = musical.format_sample_metrics(out['metrics'])
    return out


def _dp_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def train_epoch(step, X_train, len_train, batch_size, piece_size, epoch, stats, loss_accum=None, device='cuda',
                fetch_every=1, ids=None):
    """One epoch of train.py:153-200: np.random.seed(epoch), shuffle the song ids, feed every batch piece to `step`
    (= model.train_generators(...)), advance the step counter once per batch. `ids` is the PERSISTENT id array of the run
    (train.py:142 creates it once, :160 shuffles it in place every epoch, so the order is cumulative over epochs); None
    starts from arange (a single-epoch call). The reference fetches the loss after every sess.run; `fetch_every` > 1
    reads it back less often so that the host does not stall the device.
    Data parallel (torch.distributed initialised, world G > 1): every rank walks the same pieces and takes rows
    [r*b/G, (r+1)*b/G) of each piece whose row count b divides by G (other pieces are trained replicated, every rank on
    all rows: same gradient on every rank, nothing lost); the noise seed is the same on every rank because the kernels key
    their Philox streams by the GLOBAL row (`row_base`); with ragged lengths the shard's mean is re-weighted by
    (valid rows of the shard * G / valid rows of the piece) so that the allreduced gradient is the piece's."""
    import torch
    loss_accum = LossAccumulator() if loss_accum is None else loss_accum
    stats.new_epoch()
    np.random.seed(epoch)
    if ids is None:
        ids = np.arange(X_train.shape[0])
    np.random.shuffle(ids)
    loss_accum.clear()
    pending = []
    steps0 = stats.steps
    n_batches = (X_train.shape[0] + batch_size - 1) // batch_size
    rank, world = _dp_world()
    for bi, songs, len_batch in training_pieces(X_train, len_train, ids, batch_size, piece_size):
        while stats.steps < steps0 + bi:          # stats.new_step() once per finished batch (train.py:194, quirk Q11)
            stats.new_step()
        kw = {}
        len_batch = np.asarray(len_batch)
        if world > 1 and songs.shape[0] % world == 0:
            per = songs.shape[0] // world
            valid_all = float(np.minimum(len_batch, songs.shape[1]).sum())
            songs, len_batch = songs[rank * per:(rank + 1) * per], len_batch[rank * per:(rank + 1) * per]
            kw = {'row_base': rank * per, 'global_batch': per * world,
                  'loss_scale': float(np.minimum(len_batch, songs.shape[1]).sum()) * world / valid_all}
        elif world > 1:
            kw = {'row_base': 0, 'global_batch': songs.shape[0]}           # replicated piece: identical work on every rank
        loss = step(_to_device(songs, device), lengths=torch.as_tensor(len_batch), seed=stats.steps * 131 + epoch, **kw)
        pending.append(loss.detach().clone())
        if len(pending) >= fetch_every:
            for v in torch.stack(pending).flatten().cpu().tolist():
                loss_accum.update(v)
            pending = []
    if pending:
        for v in torch.stack(pending).flatten().cpu().tolist():
            loss_accum.update(v)
    while stats.steps < steps0 + n_batches:
        stats.new_step()
    return loss_accum


def fit(model, train, valid, training_config, stats=None, optimizer='adam', checkpoint_path=None, evaluate_epochs=1,
        beat_size=None, device='cuda', log=None):
    """The epoch loop of train.py:153-282 without its logging / sampling side effects: train, evaluate on the
    validation set, keep the best checkpoint (`loglik_val < stats.metric_best`, :243-257), stop after
    `early_stopping` epochs without improvement (:258-270). Returns (stats, history). Under data parallelism every rank
    runs the loop (train_epoch shards the pieces, collect_metrics the evaluation batches); rank 0 alone writes files."""
    X_train, len_train = train
    X_valid, len_valid = valid
    stats = TrainingStats() if stats is None else stats
    stats.new_run()
    batch_size = training_config['batch_size']
    if beat_size is None:
        beat_size = 1.0
    piece_size = int(training_config['piece_size'] * beat_size)
    step = model.train_generators(optimizer, training_config['learning_rate'])
    loss_accum = LossAccumulator()
    history = []
    past_epochs = stats.epoch
    loglik_val = float('inf')
    ids = np.arange(X_train.shape[0])                                    # train.py:142: one array for the whole run
    writer = _dp_world()[0] == 0
    for epoch in range(past_epochs + 1, past_epochs + training_config['epochs'] + 1):
        train_epoch(step, X_train, len_train, batch_size, piece_size, epoch, stats, loss_accum, device, ids=ids)
        rec = {'epoch': epoch, 'steps': stats.steps, 'loss': loss_accum.loss(), 'bad': loss_accum.num_bad()}
        if evaluate_epochs > 0 and epoch % evaluate_epochs == 0:
            m = collect_metrics(model, X_valid, len_valid, batch_size * 2, piece_size, device)
            loglik_val = m['log_likelihood']
            rec['valid_log_likelihood'] = loglik_val
        history.append(rec)
        if log is not None:
            log(rec)
        if loglik_val < stats.metric_best:
            stats.update_metric_best(loglik_val)
            stats.reset_idle_epochs()
            if checkpoint_path is not None and writer:
                model.save(checkpoint_path)
                stats.save(checkpoint_path + '.stats')
        else:
            stats.new_idle_epoch()
            if stats.idle_epochs >= training_config['early_stopping']:
                break
    return stats, history


def fit_encoders(model, train, valid, training_config, layer=0, stats=None, checkpoint_path=None, evaluate_epochs=1,
                 beat_size=None, device='cuda', init_songs=1600, log=None):
    """The encoder pre-training loop of train_encoders.py:17-219 for one DBN layer: run the init_ops once on
    X_train[:1600] (:106-109), then per epoch feed every batch piece to the CD-k update, evaluate the validation
    log-likelihood with the streaming mean of collect_metrics, keep the best checkpoint and stop after
    `early_stopping` idle epochs. Returns (stats, history)."""
    import torch
    X_train, len_train = train
    X_valid, len_valid = valid
    fresh = stats is None
    stats = TrainingStats() if stats is None else stats
    batch_size = training_config['batch_size']
    piece_size = int(training_config['piece_size'] * (1.0 if beat_size is None else beat_size))
    init, step = model.train_encoders(None, training_config['learning_rate'], layer=layer)
    if fresh:
        init(_to_device(X_train[:init_songs], device), lengths=torch.as_tensor(np.asarray(len_train[:init_songs])))
    stats.new_run()
    loss_accum = LossAccumulator()
    history = []
    loglik_val = float('inf')
    past_epochs = stats.epoch
    for epoch in range(past_epochs + 1, past_epochs + training_config['epochs'] + 1):
        stats.new_epoch()
        np.random.seed(epoch)
        ids = np.arange(X_train.shape[0])
        np.random.shuffle(ids)
        loss_accum.clear()
        steps0 = stats.steps
        for bi, songs, len_batch in training_pieces(X_train, len_train, ids, batch_size, piece_size):
            while stats.steps < steps0 + bi:
                stats.new_step()
            out = step(_to_device(songs, device), lengths=torch.as_tensor(np.asarray(len_batch)),
                       seed=stats.steps * 131 + epoch)
            loss_accum.update(float(out['batch/loss']))
        while stats.steps < steps0 + (X_train.shape[0] + batch_size - 1) // batch_size:
            stats.new_step()
        rec = {'epoch': epoch, 'steps': stats.steps, 'loss': loss_accum.loss()}
        if evaluate_epochs > 0 and epoch % evaluate_epochs == 0:
            tot, cnt = 0.0, 0
            for songs, seq in evaluation_pieces(X_valid, len_valid, batch_size * 2, piece_size):
                m = model.evaluate_encoders(_to_device(songs, device), lengths=torch.as_tensor(np.asarray(seq)),
                                            layer=layer, seed=epoch)
                tot += float(m['log_likelihood']) * m['rows']
                cnt += m['rows']
            loglik_val = tot / max(cnt, 1)
            rec['valid_log_likelihood'] = loglik_val
        history.append(rec)
        if log is not None:
            log(rec)
        if loglik_val < stats.metric_best:
            stats.update_metric_best(loglik_val)
            stats.reset_idle_epochs()
            if checkpoint_path is not None:
                model.save(checkpoint_path)
                stats.save(checkpoint_path + '.stats')
        else:
            stats.new_idle_epoch()
            if stats.idle_epochs >= training_config['early_stopping']:
                break
    return stats, history
