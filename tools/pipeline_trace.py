"""Debug: event timeline of one pipelined Composer training step. python tools/pipeline_trace.py [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200.common.rnn import RNN  # noqa: E402
from multinn_b200.multinn import MultINN, default_config, default_params  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = 256
model = MultINN(default_config(), default_params(mode='composer', keep_prob=0.9), 'composer')
x = torch.from_numpy((np.random.default_rng(23).random((B, T, 84, 5)) < 0.05).astype(np.uint8)).cuda()
step = model.train_generators('adam', 0.01)
for _ in range(3):
    step(x)
torch.cuda.synchronize()
RNN.TRACE = []
e0 = torch.cuda.Event(enable_timing=True)
e0.record()
step(x)
e1 = torch.cuda.Event(enable_timing=True)
e1.record()
torch.cuda.synchronize()
ev = RNN.TRACE
RNN.TRACE = None
rows = sorted(((e0.elapsed_time(e), lab) for lab, e in ev if lab), key=lambda r: r[0])
for t, lab in rows:
    print(f'{t:8.3f} ms  {lab}')
print(f'{e0.elapsed_time(e1):8.3f} ms  step end')
