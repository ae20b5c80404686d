"""Times Composer autoregressive sampling (BASELINE configs[4]: 512 steps from a 32-step intro, B sequences)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200.multinn import MultINN, default_config, default_params  # noqa: E402

B, Ti, S = int(os.environ.get('SB', 2048)), 32, int(os.environ.get('SS', 512))
model = MultINN(default_config(), default_params(mode='composer', keep_prob=0.9), 'composer')
x = torch.from_numpy((np.random.default_rng(23).random((B, Ti, 84, 5)) < 0.05).astype(np.uint8)).cuda()
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = model.generate(x, S, seed=it)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print(f'B={B} S={S}: device {e0.elapsed_time(e1):.1f} ms ({e0.elapsed_time(e1) / S * 1e3:.1f} us/step), wall {wall * 1e3:.1f} ms, '
          f'density {float(out.mean()):.4f}', flush=True)
