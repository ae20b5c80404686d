"""Debug: LSTM stack forward/backward time with and without the layer wavefront. python tools/wavefront_bench.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200.common.rnn import RNN  # noqa: E402
from multinn_b200.params import ParamArena  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = 256
arena = ParamArena()
rnn = RNN(arena, 420, [512, 256], keep_prob=0.9, name='rnn', binary_inputs=True)
arena.finalize(torch.device('cuda'), seed=1)
x = (torch.rand(T, B, 420, device='cuda') < 0.05).float()
dout = torch.randn(T, B, 256, device='cuda') * 0.01


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for mb in (0, 4096):
    RNN.WAVEFRONT_MAX_BATCH = mb
    f = timed(lambda: rnn.forward_sequence(x, keep=0.9, seed=3))

    def fb():
        rnn.forward_sequence(x, keep=0.9, seed=3)
        rnn.backward_sequence(dout)
    t = timed(fb)
    print(f'B={B} wavefront={"on" if mb else "off"}: forward {f:.2f} ms, forward+backward {t:.2f} ms (backward {t - f:.2f})')
