"""Debug: do the persistent recurrence kernels of two layers co-run on two streams? python tools/overlap_test.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = 64
bufs = []
for R in (512, 256):
    bufs.append(dict(R=R, gates=torch.randn(T, B, 4 * R, device='cuda') * 0.5, wh=torch.randn(R, 4 * R, device='cuda') * 0.05,
                     hbuf=torch.zeros(T + 1, B, R, device='cuda'), cbuf=torch.zeros(T + 1, B, R, device='cuda'),
                     out=torch.empty(T, B, R, device='cuda'), ds=torch.empty(T, B, R, device='cuda'),
                     dout=torch.randn(T, B, R, device='cuda') * 0.01, dc=torch.empty(B, R, device='cuda')))
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def fwd(b, budget):
    ops.set_sm_budget(budget)
    ops.lstm_seq_fwd(b['gates'], b['wh'], b['hbuf'], b['cbuf'], out=b['out'], dscale=b['ds'], keep=0.9, seed=1)
    ops.set_sm_budget(0)


def bwd(b, budget):
    ops.set_sm_budget(budget)
    ops.lstm_seq_bwd(b['gates'], b['wh'], b['cbuf'], b['dout'], b['ds'], b['dc'], b['dc'])
    ops.set_sm_budget(0)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def both(f, budgets):
    main = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(main)
    done = []
    for s, b, bud in zip(streams, bufs, budgets):
        with torch.cuda.stream(s):
            s.wait_event(ev)
            f(b, bud)
            d = torch.cuda.Event()
            d.record(s)
            done.append(d)
    for d in done:
        main.wait_event(d)


for name, f, budgets in (('fwd', fwd, (96, 48)), ('bwd', bwd, (72, 36))):
    for _ in range(2):
        a = timed(lambda: f(bufs[0], 0))
        b = timed(lambda: f(bufs[1], 0))
        ab = timed(lambda: f(bufs[0], budgets[0]))
        bb = timed(lambda: f(bufs[1], budgets[1]))
        c = timed(lambda: both(f, budgets))
    print(f'B={B} {name}: L1 {a:.2f} ms, L2 {b:.2f} ms, with budgets {ab:.2f} / {bb:.2f}, both streams {c:.2f} ms (T={T})')
