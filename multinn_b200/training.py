"""Optimiser step and data-parallel exchange (mirrors reference utils/training.py:151-177 compute_gradients and
train.py:61-64): grads of `batch/loss` over the listed variables -> clip_by_global_norm(5.) -> TF-style Adam
(epsilon on the uncorrected sqrt(v)) or SGD. Data parallel: ONE NCCL allreduce of the flat fp32 gradient
buffer per step, before the clip (the clip must see the global-batch gradient)."""
import torch
import torch.distributed as dist

from . import ops


class AdamOptimizer:
    """tf.train.AdamOptimizer(learning_rate, epsilon=1e-4) as constructed at train.py:64."""

    def __init__(self, learning_rate=0.01, beta1=0.9, beta2=0.999, epsilon=1e-4):
        self.lr, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon
        self.kind = 'adam'


class GradientDescentOptimizer:
    """tf.train.GradientDescentOptimizer (train.py:61-62, --sgd)."""

    def __init__(self, learning_rate=0.01):
        self.lr = learning_rate
        self.kind = 'sgd'


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def dp_row_map(batch_local, row_base=None, global_batch=None):
    """ops.row_map arguments of a data-parallel shard: this rank holds sequences [row_base, row_base + batch_local) of a
    global batch (defaults: equal shards in rank order). The kernels key their Philox streams (dropout, DBN
    codes, Gibbs chains, NADE sampling) by the GLOBAL sequence index, so every rank uses the SAME seed and the noise --
    hence the loss and the samples -- does not depend on the number of GPUs."""
    r, ws = world()
    return (batch_local, batch_local * ws if global_batch is None else global_batch,
            r * batch_local if row_base is None else row_base)


def shard_batch(x, rank=None, world_size=None):
    """Data-parallel partition of a global batch along dim 0: rank r gets rows [r*B/G, (r+1)*B/G).
    Equal shards are required so that the mean of per-rank mean-loss gradients is the global-batch gradient."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    B = x.shape[0]
    if B % world_size:
        raise ValueError(f'global batch {B} is not divisible by the world size {world_size}')
    per = B // world_size
    return x[rank * per:(rank + 1) * per]


def allreduce_sum_(flat):
    """The one exchange step of a training step: in-place sum-allreduce of the flat gradient bucket (NCCL on GPUs,
    gloo in the CPU tests). Returns the factor that turns the sum of per-rank means into the global mean."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(flat)
    return 1.0 / ws


class GradientApplier:
    """compute_gradients(): allreduce -> global norm -> clip -> apply, over one contiguous arena range."""

    def __init__(self, arena, optimizer, lr=None, clip_norm=5.0, offset=0, length=None):
        self.arena, self.opt = arena, optimizer
        self.lr = optimizer.lr if lr is None else lr
        self.clip_norm = clip_norm                      # hard-coded 5. in the reference (training.py:166, quirk Q13)
        self.offset = offset
        self.length = arena.size - offset if length is None else length
        self.step_count = 0
        self.sqnorm = torch.zeros(1, device=arena.flat.device)
        if optimizer.kind == 'adam':
            arena.ensure_slots()

    def _rng(self, t):
        return t[self.offset:self.offset + self.length]

    def zero_grad(self):
        self._rng(self.arena.grad).zero_()

    def apply(self):
        g = self._rng(self.arena.grad)
        grad_scale = allreduce_sum_(g)                  # sum of per-rank mean-loss grads; averaged in the optimiser kernel
        ops.sqnorm_into(g, self.sqnorm)
        self.step_count += 1
        p = self._rng(self.arena.flat)
        if self.opt.kind == 'adam':
            ops.clip_adam(p, g, self._rng(self.arena.m), self._rng(self.arena.v), self.sqnorm, self.step_count,
                          self.lr, grad_scale=grad_scale, clip_norm=self.clip_norm, beta1=self.opt.beta1,
                          beta2=self.opt.beta2, eps=self.opt.epsilon)
        else:
            ops.clip_sgd(p, g, self.sqnorm, self.lr, grad_scale=grad_scale, clip_norm=self.clip_norm)

    def grad_norm(self):
        """Global gradient norm of the last step (after averaging over ranks)."""
        _, ws = world()
        return float(self.sqnorm.sqrt()) / ws


class BatchPrefetcher:
    """Overlaps the host->device copy of the NEXT batch with the current training step: `put(host)` enqueues the copy
    on a side stream into one of two device buffers, `get()` makes the current stream wait for it and returns the
    buffer. The reference feeds NumPy batches through feed_dict each sess.run (train.py:186-189); this is the
    equivalent input path for pinned host batches (bool/uint8 or float32 piano-rolls)."""

    def __init__(self, device='cuda'):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.bufs = [None, None]
        self.ready = [None, None]       # copy finished (side stream)
        self.released = [None, None]    # last consumer enqueued (main stream)
        self.head = 0                   # next slot to fill
        self.tail = 0                   # next slot to hand out
        self.pending = 0

    def put(self, host):
        if self.pending >= 2:
            raise RuntimeError('BatchPrefetcher holds two batches already: call get() first')
        i = self.head
        if self.bufs[i] is None or self.bufs[i].shape != host.shape or self.bufs[i].dtype != host.dtype:
            self.bufs[i] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
        if self.released[i] is not None:
            self.stream.wait_event(self.released[i])        # the step that read this buffer has been enqueued and run
        with torch.cuda.stream(self.stream):
            self.bufs[i].copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ready[i] = ev
        self.head ^= 1
        self.pending += 1

    def get(self):
        if self.pending == 0:
            raise RuntimeError('BatchPrefetcher is empty: call put() first')
        i = self.tail
        torch.cuda.current_stream().wait_event(self.ready[i])
        self.tail ^= 1
        self.pending -= 1
        self._last = i
        return self.bufs[i]

    def release(self):
        """Call after the step that consumes the last `get()` buffer has been enqueued."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.released[self._last] = ev
