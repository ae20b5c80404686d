"""CPU, world_size 2 over gloo: the host-side data-parallel logic (batch sharding, the single flat-bucket allreduce and
its mean factor). The CUDA kernels are not involved; the same functions run under NCCL on the GPUs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from multinn_b200.training import allreduce_sum_, dp_row_map, shard_batch, world as w
    assert w() == (rank, world)
    assert dp_row_map(4) == (4, 4 * world, 4 * rank)      # noise is keyed by the GLOBAL sequence index, same seed everywhere
    x = torch.arange(8 * 3, dtype=torch.float32).view(8, 3)
    shard = shard_batch(x)
    assert shard.shape[0] == 4 and float(shard[0, 0]) == rank * 12
    # per-rank "mean-loss gradient" of its shard; sum * factor must equal the global-batch mean gradient
    g = shard.mean(0).clone()
    scale = allreduce_sum_(g)
    q.put((rank, (g * scale).numpy(), scale))
    try:
        shard_batch(torch.zeros(7, 2))
    except ValueError:
        q.put((rank, 'uneven-rejected', 0))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_gives_global_mean_gradient():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(4)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = torch.arange(24, dtype=torch.float32).view(8, 3).mean(0).numpy()
    vals = [g for g in got if not isinstance(g[1], str)]
    assert len(vals) == 2 and all(abs(g[2] - 0.5) < 1e-12 for g in vals)
    for _, v, _ in vals:
        np.testing.assert_allclose(v, ref, rtol=1e-6)
    assert sum(1 for g in got if isinstance(g[1], str)) == 2


class _EvalStub:
    """evaluate() whose per-row NLL is a deterministic function of the piece's content (no device needed)."""

    def evaluate(self, x, lengths=None):
        per_row = x.float().sum((2, 3))                                   # [B, T]
        rows = [per_row[b, :int(l)] for b, l in enumerate(lengths)]
        return {'nll': torch.cat(rows)[:, None].repeat(1, 2) * 0.01}


def _eval_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from multinn_b200.utils import training as U
    X, lengths = _eval_data()
    q.put((rank, U.collect_metrics(_EvalStub(), X, lengths, batch_size=3, piece_size=4, device='cpu')))
    dist.destroy_process_group()


def _eval_data():
    rng = np.random.default_rng(11)
    return (rng.random((10, 9, 6, 2)) < 0.3).astype(np.uint8), np.array([9, 9, 4, 7, 9, 2, 9, 9, 5, 9])


def test_streaming_evaluation_is_sharded_and_allreduced():
    """collect_metrics under world_size 2: each rank evaluates every other batch, one allreduce of (sum, sum exp, count);
    both ranks report the single-process metrics of the whole dataset."""
    from multinn_b200.utils import training as U
    X, lengths = _eval_data()
    ref = U.collect_metrics(_EvalStub(), X, lengths, batch_size=3, piece_size=4, device='cpu')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in got) == [0, 1]
    for _, m in got:
        assert m['rows'] == ref['rows'] and m['rows'] > 0
        assert abs(m['log_likelihood'] - ref['log_likelihood']) < 1e-12
        assert abs(m['perplexity'] - ref['perplexity']) < 1e-12
