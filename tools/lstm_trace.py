"""Debug: per-phase globaltimer trace of the pair forward recurrence (MNN_LSTM_TRACE=<file> must be set).
Runs one layer-1-shaped and one layer-2-shaped forward at B=2048, T=64 and prints per-step phase offsets."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200 import ops  # noqa: E402

T, B = 64, int(os.environ.get('TB', 2048))
for R in (512, 256):
    gates = torch.randn(T, B, 4 * R, device='cuda') * 0.5
    wh = torch.randn(R, 4 * R, device='cuda') * 0.05
    hbuf = torch.zeros(T + 1, B, R, device='cuda')
    cbuf = torch.zeros(T + 1, B, R, device='cuda')
    out = torch.empty(T, B, R, device='cuda')
    ds = torch.empty(T, B, R, device='cuda')
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.lstm_seq_fwd(gates, wh, hbuf, cbuf, out=out, dscale=ds, keep=0.9, seed=1)
        e1.record()
        torch.cuda.synchronize()
    print(f'R={R}: fwd {e0.elapsed_time(e1) / T * 1e3:.1f} us/step')
    dout = torch.randn(T, B, R, device='cuda') * 0.01
    dc = torch.empty(B, R, device='cuda')
    for _ in range(2):
        g2 = gates.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.lstm_seq_bwd(g2, wh, cbuf, dout, ds, dc, dc)
        e1.record()
        torch.cuda.synchronize()
    print(f'R={R}: bwd {e0.elapsed_time(e1) / T * 1e3:.1f} us/step')
path = os.environ.get('MNN_LSTM_TRACE')
if path and os.path.exists(path):
    FW = ['flag_ok', 'tma_issued', 'conv_first_full', 'conv_done', 'mma_first', 'mma_commit', 'epi0_tfull',
          'epi0_done', 'published', 'epi1_tfull', 'epi1_done', 'epi0_tiles_in']
    BW = ['cntB_ok', 'mma_first', 'mma_commit', 'epiA_tfull', 'cntA_inc', 'cell_cntA_ok', 'cell_tiles_in', 'cell_done',
          'cell_cntB_inc', 'cells_all_done']
    blocks = open(path).read().split('# lstm pair ')[1:]
    for blk in blocks[1::2]:          # second (warm) call of each shape
        lines = blk.strip().split('\n')
        print('#', lines[0])
        rows = np.array([[int(v) for v in ln.split()] for ln in lines[1:]], dtype=np.int64)
        ev = rows[:, 2:].astype(np.float64)
        ev[ev == 0] = np.nan
        for st in (2, 3):
            sel = rows[:, 1] == st
            e = ev[sel]
            t0 = np.nanmin(e[:, 0])
            names = FW if lines[0].startswith('fwd') else BW
            print(f' step {st}: ' + ' | '.join(f'{n} {np.nanmin(e[:, i]) - t0:.0f}/{np.nanmean(e[:, i]) - t0:.0f}/{np.nanmax(e[:, i]) - t0:.0f}'
                                                for i, n in enumerate(names)))
        # step period
        per = [np.nanmin(ev[rows[:, 1] == st + 1][:, 0]) - np.nanmin(ev[rows[:, 1] == st][:, 0]) for st in range(1, 6)]
        print(' step period ns:', per)
