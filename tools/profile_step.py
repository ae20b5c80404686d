"""Two Composer training steps at a small batch: the command profiled with ncu (see profiles/)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200.multinn import MultINN, default_config, default_params  # noqa: E402

B, T = int(os.environ.get('PB', 512)), int(os.environ.get('PT', 64))
model = MultINN(default_config(), default_params(mode='composer', keep_prob=0.9), 'composer')
step = model.train_generators('adam', 0.01)
x = torch.from_numpy((np.random.default_rng(23).random((B, T, 84, 5)) < 0.05).astype(np.float32)).cuda()
for _ in range(int(os.environ.get('PSTEPS', 2))):
    loss = step(x)
torch.cuda.synchronize()
print('loss', float(loss))
