"""Launches the kernels the bench step does not cover (or covers at one shape only) a few times each, for ONE
`ncu --set full` capture: the fused Gibbs chain at generation and training row counts, the NADE sampler, the weight-resident
LSTM recurrences at the 8-GPU shard shape, the tcgen05 NADE forward next to the SIMT forward / backward at C5 rows.
  python tools/profile_misc.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200 import ops  # noqa: E402

g = torch.Generator(device='cuda').manual_seed(0)
rnd = lambda *s: torch.randn(*s, device='cuda', generator=g)

# ---- RBM Gibbs chain (C3 generator RBM 84 x 256, k = 10)
W = rnd(84, 256) * 0.1
for N in (2048, 65536):
    v0 = (torch.rand(N, 84, device='cuda', generator=g) < 0.05).float()
    bh, bv = rnd(N, 256) * 0.3, rnd(N, 84) * 0.3 - 2
    vk = torch.empty(N, 84, device='cuda')
    for _ in range(2):
        ops.rbm_gibbs(v0, W, bh, bv, 10, v_k=vk, seed=1, offset=0)

# ---- NADE sampler + forward + backward at C5's generation batch / training rows
M, D, H = 5, 84, 256
we, wd = rnd(M, D, H) / D ** 0.5, rnd(M, D, H) / D ** 0.5
B = 2048
fc = rnd(B, M * (H + D))
fc[:, M * H:] -= 3.0
out = torch.empty(B, M * D, device='cuda')
for _ in range(2):
    ops.nade_sample(fc, 0, M * H, we, wd, out, M * D, M, 1, use_philox=True, seed=3, offset=0)
N = 131072
fcN = rnd(N, M * (H + D))
fcN[:, M * H:] -= 2.0
x = (torch.rand(M, N, D, device='cuda', generator=g) < 0.05).float()
bits = torch.empty(M, N, 4, dtype=torch.int32, device='cuda')
for m in range(M):
    ops.pack_rows(x[m], bits[m], D)
nll, dfc = torch.empty(M, N, device='cuda'), torch.zeros_like(fcN)
dwe, dwd = torch.zeros_like(we), torch.zeros_like(wd)
for mode in ('simt', 'tc'):
    ops.set_nade_mode(mode)
    for _ in range(2):
        ops.nade_logprob_fwd(bits, fcN, 0, M * H, we, wd, nll, dfc=dfc, gscale=1.0 / (N * M))
ops.set_nade_mode('simt')
for _ in range(2):
    ops.nade_logprob_bwd(bits, fcN, 0, M * H, we, wd, dfc, dwe, dwd)

# ---- weight-resident recurrences at the 8-GPU shard (B = 256), both layers, 64 steps
T, Bq = 64, 256
for R in (512, 256):
    gates = rnd(T, Bq, 4 * R)
    wh = rnd(R, 4 * R) * 0.05
    hbuf, cbuf = torch.zeros(T + 1, Bq, R, device='cuda'), torch.zeros(T + 1, Bq, R, device='cuda')
    o, ds = torch.empty(T, Bq, R, device='cuda'), torch.empty(T, Bq, R, device='cuda')
    dout = rnd(T, Bq, R)
    dh, dc = torch.empty(Bq, R, device='cuda'), torch.empty(Bq, R, device='cuda')
    for _ in range(2):
        gg = gates.clone()
        ops.lstm_seq_fwd(gg, wh, hbuf, cbuf, out=o, dscale=ds, keep=0.9, seed=1, mode='tc', persistent=True)
        if os.environ.get('MISC_SKIP_BWD') != '1':     # ncu cannot replay the cluster + cooperative BPTT kernel (r2o: error 9)
            ops.lstm_seq_bwd(gg, wh, cbuf, dout, ds, dh, dc, mode='tc', persistent=True)
torch.cuda.synchronize()
print('profile_misc ok')
