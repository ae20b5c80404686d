import sys, time, torch
sys.path.insert(0, '/root/repo')
from multinn_b200 import ops
from multinn_b200.common.rbm import RBM
from multinn_b200.params import ParamArena
arena = ParamArena(); rbm = RBM(84, 256, k=10, arena=arena, name='rbm'); arena.finalize('cuda', seed=3)
g = torch.Generator(device='cuda').manual_seed(0)
for N in (256, 1024, 2048, 4096, 8192, 2048):
    v = (torch.rand(N, 84, device='cuda', generator=g) < 0.05).float()
    bh = torch.randn(N, 256, device='cuda', generator=g) * 0.3; bv = torch.randn(N, 84, device='cuda', generator=g) * 0.3 - 2
    vk = torch.empty(N, 84, device='cuda'); pv = torch.empty(N, 84, device='cuda')
    for label, fn in (('class', lambda: rbm.sample(v, bh, bv)),
                      ('ops', lambda: ops.rbm_gibbs(v, rbm.W.data, bh, bv, 10, p_v=pv, v_k=vk, seed=5, offset=0)),
                      ('ops_bigseed', lambda: ops.rbm_gibbs(v, rbm.W.data, bh, bv, 10, p_v=pv, v_k=vk, seed=rbm._key(0, 1), offset=12345))):
        for _ in range(2): fn()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print(N, label, round(e0.elapsed_time(e1) / 10, 4), 'ms dev', round((time.perf_counter() - t0) * 100, 4), 'ms wall', flush=True)
