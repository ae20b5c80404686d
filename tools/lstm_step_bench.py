"""Per-step time of the forward recurrence at small per-GPU batches: weight-resident kernel (lstm_res.cu) vs the
streaming 1-CTA kernel (MNN_LSTM_RES=0 in a second process). python tools/lstm_step_bench.py [B] [R] [T]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
R = int(sys.argv[2]) if len(sys.argv) > 2 else 512
T = int(sys.argv[3]) if len(sys.argv) > 3 else 256
g = torch.Generator(device='cuda').manual_seed(0)
gates0 = torch.randn(T, B, 4 * R, device='cuda', generator=g)
wh = torch.randn(R, 4 * R, device='cuda', generator=g) * 0.05
hbuf, cbuf = torch.zeros(T + 1, B, R, device='cuda'), torch.zeros(T + 1, B, R, device='cuda')
out, dscale = torch.empty(T, B, R, device='cuda'), torch.empty(T, B, R, device='cuda')
ops.set_sm_budget(int(os.environ.get('SMB', 0)))
res = []
for it in range(4):
    gates = gates0.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.lstm_seq_fwd(gates, wh, hbuf, cbuf, out=out, dscale=dscale, keep=0.9, seed=1, mode='tc', persistent=True)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1))
dout = torch.randn(T, B, R, device='cuda', generator=g)
dh_work, dc_work = torch.empty(B, R, device='cuda'), torch.empty(B, R, device='cuda')
resb = []
gact = gates.clone()
for it in range(4):
    gb = gact.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.lstm_seq_bwd(gb, wh, cbuf, dout, dscale, dh_work, dc_work, mode='tc', persistent=True)
    e1.record()
    torch.cuda.synchronize()
    resb.append(e0.elapsed_time(e1))
print(f'  bwd: {min(resb):.3f} ms = {min(resb) / T * 1e3:.2f} us/step; checksum {float(gb.double().sum()):.5f} '
      f'{float(gb.double().abs().sum()):.3f} RES_BWD={os.environ.get("MNN_LSTM_RES_BWD", "1")}', flush=True)
print(f'B={B} R={R} T={T} MNN_LSTM_RES={os.environ.get("MNN_LSTM_RES", "1")} SMB={os.environ.get("SMB", 0)}: {min(res):.3f} ms = {min(res) / T * 1e3:.2f} us/step; '
      f'checksum {float(hbuf[T].double().sum()):.6f} {float(out.double().sum()):.6f}', flush=True)
