"""Data-parallel consistency check: G ranks each train on their shard of a global batch (one NCCL allreduce of the flat
gradient bucket per step); rank 0 also trains a private copy on the WHOLE batch. Weights must agree -- with dropout
off (keep 1.0) AND with in-kernel Philox dropout (keep 0.9): the noise is keyed by the global batch row (ops.row_map),
so the sharded run draws the masks of the single-GPU run.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200.multinn import MultINN, default_config, default_params  # noqa: E402
from multinn_b200.training import shard_batch  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for keep in (1.0, 0.9):
        mk = lambda: MultINN(default_config(), default_params(mode='composer', num_hidden=128, num_hidden_rnn=(64, 32),
                                                              keep_prob=keep), 'composer')
        model = mk()
        step = model.train_generators('adam', 0.01)
        rng = np.random.default_rng(0)
        xs = [torch.from_numpy((rng.random((8 * world, 12, 84, 5)) < 0.08).astype(np.float32)).cuda() for _ in range(4)]
        losses = []
        for x in xs:
            l = step(shard_batch(x))
            dist.all_reduce(l)
            losses.append(float(l) / world)
        if rank == 0:
            # single-process reference on the whole batch: detach from the process group by monkeypatching world()
            import multinn_b200.training as tr
            ref = mk()
            saved = tr.world
            tr.world = lambda: (0, 1)
            try:
                rstep = ref.train_generators('adam', 0.01)
                rl = [float(rstep(x)) for x in xs]
            finally:
                tr.world = saved
            dw = float((ref.arena.flat - model.arena.flat).abs().max())
            dl = max(abs(a - b) / abs(b) for a, b in zip(losses, rl))
            print(f'dp_check world={world} keep={keep}: max |dW| = {dw:.3e}, max rel loss diff = {dl:.3e}')
            ok = ok and dw < 2e-4 and dl < 1e-5
    flag = torch.tensor([1.0 if ok else 0.0], device='cuda')
    dist.broadcast(flag, 0)
    # every rank must hold identical weights after identical updates
    w0 = model.arena.flat.clone()
    dist.broadcast(w0, 0)
    same = bool(torch.equal(w0, model.arena.flat))
    dist.destroy_process_group()
    if not (float(flag) > 0 and same):
        sys.exit(1)


if __name__ == '__main__':
    main()
