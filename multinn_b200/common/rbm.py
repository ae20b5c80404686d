"""Binary RBM (mirrors reference models/common/rbm.py:17-387).

Bernoulli sampling contract (TFP 0.6.0, rbm.py:375-387): sample = float(u < p), strict. Every sampling method takes
optional uniforms `u` (parity runs) and otherwise uses the in-kernel Philox generator keyed by (seed, offset).
The k-step Gibbs chain runs as ONE launch of the fused kernel `mnn_rbm_gibbs` (W and its transpose in shared memory,
bias + sigmoid + Bernoulli in registers, csrc/rbm.cu) when the shape fits (D, H <= 256, e.g. the 84 x 256 generator RBM
and the 84/168 DBN layers) and the batch is below the measured crossover (`ops.GIBBS_MODE = "auto"`: <= 32 768 rows,
i.e. generation and small training batches); larger batches, larger layers (the Joint encoder's 420 x 168) and single
half-steps (`forward`, `reconstruct`) run a tensor-core GEMM followed by the fused bias + sigmoid + Bernoulli kernel.
"""
import zlib

import torch

from .. import ops
from ..params import glorot_uniform, zeros
from .model import Model


class RBM(Model):
    def __init__(self, num_dims, num_hidden, k=1, name='rbm', arena=None):
        super().__init__(name=name)
        self._num_dims, self._num_hidden, self._k = num_dims, num_hidden, k
        self.W = arena.add(f'{name}/W', (num_dims, num_hidden), glorot_uniform(num_dims, num_hidden))   # rbm.py:36-44
        self.bh = arena.add(f'{name}/bh', (1, num_hidden), zeros())                                      # :47-52
        self.bv = arena.add(f'{name}/bv', (1, num_dims), zeros())                                        # :53-58
        self._seed = 0
        self._calls = 0
        # Every RBM draws from its own Philox stream: the reference's Bernoulli draws of distinct RBMs (DBN layers, the
        # per-track encoders, the generator's RBM) are independent (tfp Bernoulli.sample at rbm.py:375-387 has one op-level
        # seed per graph node), so the kernels' key = seed is mixed with a stream id derived from the variable scope name,
        # and with a domain tag that separates the fused chain (counter = row) from the half-step kernels (counter = element).
        self._stream = zlib.crc32(name.encode())

    num_dims = property(lambda s: s._num_dims)
    num_hidden = property(lambda s: s._num_hidden)
    k = property(lambda s: s._k)

    @property
    def trainable_params(self):
        return [self.W, self.bh, self.bv]

    def _key(self, seed, domain=0):
        """64-bit Philox key of this RBM for a user seed: splitmix-style mix of (seed, stream id, domain)."""
        z = ((self._seed if seed is None else int(seed)) + 0x9E3779B97F4A7C15 * (2 * self._stream + domain + 1)) % (1 << 64)
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) % (1 << 64)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) % (1 << 64)
        return z ^ (z >> 31)

    def _next_offset(self, n):
        o = self._calls
        self._calls += n * ops.row_scale()      # counters advance by GLOBAL element counts (ops.row_map)
        return o

    # ------------------------------------------------------------------ conditionals (rbm.py:337-373)
    def _cond_prob_h(self, v, bh=None):
        bh = self.bh.data if bh is None else bh
        pre = torch.empty(v.shape[0], self._num_hidden, device=v.device)
        ops.gemm(v, self.W.data, pre)
        return pre, bh

    def _cond_prob_v(self, h, bv=None):
        bv = self.bv.data if bv is None else bv
        pre = torch.empty(h.shape[0], self._num_dims, device=h.device)
        ops.gemm(h, self.W.data, pre, transB=True)
        return pre, bv

    def forward(self, v, bh=None, u=None, seed=None, sample=True):
        """rbm.py:148-167: p_h = sigmoid(v W + bh), h ~ Bernoulli(p_h). Returns (p_h, h)."""
        pre, bh = self._cond_prob_h(v, bh)
        p = torch.empty_like(pre)
        h = torch.empty_like(pre) if sample else None
        ops.bias_sigmoid_sample(pre, bias=bh, u=u, p=p, s=h, use_philox=sample and u is None,
                                seed=self._key(seed), offset=self._next_offset(pre.numel()))
        return p, h

    def reconstruct(self, h, bv=None, u=None, seed=None, sample=True):
        """rbm.py:169-190: p_v = sigmoid(h W^T + bv), v ~ Bernoulli(p_v). Returns (p_v, v)."""
        pre, bv = self._cond_prob_v(h, bv)
        p = torch.empty_like(pre)
        v = torch.empty_like(pre) if sample else None
        ops.bias_sigmoid_sample(pre, bias=bv, u=u, p=p, s=v, use_philox=sample and u is None,
                                seed=self._key(seed), offset=self._next_offset(pre.numel()))
        return p, v

    def sample(self, v, bh=None, bv=None, k=None, u=None, seed=None):
        """rbm.py:192-231: k-step Gibbs chain from v. k=None -> self.k (quirk Q1: the docstring's intent).
        u = (uh[k,N,H], uv[k,N,D]) or None. Returns (p_v of the last step, v_k); k == 0 returns (v, v)."""
        k = self._k if k is None else k
        if k > 0 and ops.gibbs_use_fused(v.shape[0]):
            bh_ = self.bh.data if bh is None else bh
            bv_ = self.bv.data if bv is None else bv
            if ops.rbm_gibbs_supported(v, self.W.data, bh_, bv_, u):
                N = v.shape[0]
                p_v = torch.empty(N, self._num_dims, device=v.device)
                vk = torch.empty(N, self._num_dims, device=v.device)
                uu = None if u is None else (u[0][:k], u[1][:k])
                # Philox: the chain's stream is keyed by the row, so only one offset unit per call is consumed per row
                ops.rbm_gibbs(v, self.W.data, bh_, bv_, k, p_v=p_v, v_k=vk, u=uu,
                              seed=self._key(seed, domain=1), offset=self._next_offset(N))
                return p_v, vk
        p_v, vk = v, v
        for s in range(k):
            _, hk = self.forward(vk, bh, u=None if u is None else u[0][s], seed=seed)
            p_v, vk = self.reconstruct(hk, bv, u=None if u is None else u[1][s], seed=seed)
        return p_v, vk

    # ------------------------------------------------------------------ free energy (rbm.py:233-263)
    def free_energy(self, v, bh=None, bv=None):
        """F(v)[n] = -sum_j log(1 + exp((v W + bh)_j)) - v . bv (the un-broadcast [N] vector, quirk Q4)."""
        pre, bh = self._cond_prob_h(v, bh)
        bv = self.bv.data if bv is None else bv
        F = torch.empty(v.shape[0], device=v.device)
        ops.rbm_free_energy(pre, bh, v, bv, F)
        return F

    def free_energy_cost(self, v, v_sample, bh=None, bv=None):
        """(mean cost, mean free energy): reduce_mean over the reference's [N,N] broadcast equals
        mean(F(v)) - mean(F(v_sample)) (metrics/statistical.py:34, quirk Q4)."""
        Fv, Fs = self.free_energy(v, bh, bv), self.free_energy(v_sample, bh, bv)
        out = torch.zeros(2, device=v.device)
        ops.sum_into(Fv, out[0:1], scale=1.0 / v.shape[0])
        ops.sum_into(Fs, out[1:2], scale=1.0 / v.shape[0])
        return out[0:1] - out[1:2], out[0:1]

    def free_energy_cost_backward(self, v, v_sample, scale=1.0):
        """Gradient of scale * (mean F(v) - mean F(v_sample)) wrt W, bh, bv (internal biases, quirk Q3), ACCUMULATED
        into the arena grads: dF/dW = -v^T sigmoid(v W + bh), dF/dbh = -sigmoid(.), dF/dbv = -v."""
        N = v.shape[0]
        a = scale / N
        for vv, sgn in ((v, -a), (v_sample, a)):
            pre, bh = self._cond_prob_h(vv)
            p = torch.empty_like(pre)
            ops.bias_sigmoid_sample(pre, bias=bh, p=p)
            ops.gemm(vv, p, self.W.grad, transA=True, alpha=sgn, beta=1.0)
            tmp = torch.empty(self._num_hidden, device=v.device)
            ops.colsum(p, tmp)
            ops.axpy(self.bh.grad.view(-1), tmp, sgn)
            tmpv = torch.empty(self._num_dims, device=v.device)
            ops.colsum(vv, tmpv)
            ops.axpy(self.bv.grad.view(-1), tmpv, sgn)

    # ------------------------------------------------------------------ CD-k (rbm.py:265-335)
    def visible_bias_init(self, v):
        """rbm.py:286-297: bv = log(1e-6 + p / (1 - p)), p = mean(v, 0)."""
        p = torch.empty(self._num_dims, device=v.device)
        ops.colsum(v, p)
        p = p / v.shape[0]
        self.bv.data.copy_(torch.log(1e-6 + p / (1 - p)).view(1, -1))

    def train(self, v, lr, u=None, seed=None):
        """CD-k update (rbm.py:299-335), applied in place (assign_add, no optimiser, no clipping):
        dW = lr/N (v^T h - p_vk^T p_hk), dbv = lr/N sum(v - p_vk), dbh = lr/N sum(h - p_hk), with h ~ p(h|v) SAMPLED and
        p_hk = p(h|v_k). u = dict(uh[k,N,H], uv[k,N,D], uh0[N,H], uhk[N,H]) or None. Data parallel: the three
        sufficient statistics are summed over ranks and N is the global row count."""
        import torch.distributed as dist
        N = v.shape[0]
        p_vs, vs = self.sample(v, k=self._k, u=None if u is None else (u['uh'], u['uv']), seed=seed)
        _, h = self.forward(v, u=None if u is None else u['uh0'], seed=seed)
        p_hs, _ = self.forward(vs, u=None if u is None else u['uhk'], seed=seed)
        D, H = self._num_dims, self._num_hidden
        stats = torch.empty(D * H + D + H, device=v.device)
        dW = stats[:D * H].view(D, H)
        ops.gemm(v, h, dW, transA=True)
        ops.gemm(p_vs, p_hs, dW, transA=True, alpha=-1.0, beta=1.0)
        tmp = torch.empty(max(D, H), device=v.device)
        ops.colsum(v, stats[D * H:D * H + D])
        ops.colsum(p_vs, tmp[:D])
        ops.axpy(stats[D * H:D * H + D], tmp[:D], -1.0)
        ops.colsum(h, stats[D * H + D:])
        ops.colsum(p_hs, tmp[:H])
        ops.axpy(stats[D * H + D:], tmp[:H], -1.0)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(stats)
            N = N * dist.get_world_size()
        a = lr / N
        ops.axpy(self.W.data.view(-1), stats[:D * H], a)
        ops.axpy(self.bv.data.view(-1), stats[D * H:D * H + D], a)
        ops.axpy(self.bh.data.view(-1), stats[D * H + D:], a)
        return p_vs, vs
