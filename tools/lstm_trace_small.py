"""Debug: per-phase globaltimer trace of the recurrence kernels at a small per-GPU batch (MNN_LSTM_TRACE=<file> must be
set): the 1-CTA forward kernel and the pair BPTT kernel under the wavefront's SM budgets. python tools/lstm_trace_small.py [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = 64
for R, fb, bb in ((512, 96, 72), (256, 48, 36)):
    gates = torch.randn(T, B, 4 * R, device='cuda') * 0.5
    wh = torch.randn(R, 4 * R, device='cuda') * 0.05
    hbuf = torch.zeros(T + 1, B, R, device='cuda')
    cbuf = torch.zeros(T + 1, B, R, device='cuda')
    out = torch.empty(T, B, R, device='cuda')
    ds = torch.empty(T, B, R, device='cuda')
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.set_sm_budget(fb)
        e0.record()
        ops.lstm_seq_fwd(gates, wh, hbuf, cbuf, out=out, dscale=ds, keep=0.9, seed=1)
        e1.record()
        ops.set_sm_budget(0)
        torch.cuda.synchronize()
    print(f'B={B} R={R}: fwd {e0.elapsed_time(e1) / T * 1e3:.1f} us/step')
    dout = torch.randn(T, B, R, device='cuda') * 0.01
    dc = torch.empty(B, R, device='cuda')
    for _ in range(2):
        g2 = gates.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.set_sm_budget(bb)
        e0.record()
        ops.lstm_seq_bwd(g2, wh, cbuf, dout, ds, dc, dc)
        e1.record()
        ops.set_sm_budget(0)
        torch.cuda.synchronize()
    print(f'B={B} R={R}: bwd {e0.elapsed_time(e1) / T * 1e3:.1f} us/step')
path = os.environ.get('MNN_LSTM_TRACE')
if path and os.path.exists(path):
    F1 = ['flag_ok', 'tma_issued', 'conv_first_full', 'conv_last_full', 'mma_first', 'mma_commit', 'epi_tfull', 'epi_done',
          'published']
    FW = ['flag_ok', 'tma_issued', 'conv_first_full', 'conv_done', 'mma_first', 'mma_commit', 'epi0_tfull',
          'epi0_done', 'published', 'epi1_tfull', 'epi1_done', 'epi0_tiles_in']
    BW = ['cntB_ok', 'mma_first', 'mma_commit', 'epiA_tfull', 'cntA_inc', 'cell_cntA_ok', 'cell_tiles_in', 'cell_done',
          'cell_cntB_inc', 'cells_all_done']
    blocks = open(path).read().split('# lstm ')[1:]
    seen = {}
    for blk in blocks:
        lines = blk.strip().split('\n')
        seen[lines[0]] = lines            # keep the last (warm) call of each header
    for head, lines in seen.items():
        print('#', head)
        rows = np.array([[int(v) for v in ln.split()] for ln in lines[1:]], dtype=np.int64)
        ev = rows[:, 2:].astype(np.float64)
        ev[ev == 0] = np.nan
        names = F1 if head.startswith('1cta') else (FW if 'fwd' in head else BW)
        for st in (2, 3):
            e = ev[rows[:, 1] == st]
            t0 = np.nanmin(e[:, 0])
            print(f' step {st}: ' + ' | '.join(f'{n} {np.nanmin(e[:, i]) - t0:.0f}/{np.nanmean(e[:, i]) - t0:.0f}/{np.nanmax(e[:, i]) - t0:.0f}'
                                                for i, n in enumerate(names)))
        per = [np.nanmin(ev[rows[:, 1] == st + 1][:, 0]) - np.nanmin(ev[rows[:, 1] == st][:, 0]) for st in range(1, 6)]
        print(' step period ns:', per)
