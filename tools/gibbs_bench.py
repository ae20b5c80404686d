"""Times the k-step Gibbs chain at config C3's generator shape (N = 512*128 rows, 84 x 256, k = 10, per-row biases,
Philox) in both modes: "fused" (mnn_rbm_gibbs, one launch) and "gemm" (2k GEMM + half-step launches).
Usage (GPU box): python tools/gibbs_bench.py [N]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(N=65536, D=84, H=256, k=10, iters=10):
    import torch
    from multinn_b200 import _lib, ops
    from multinn_b200.common.rbm import RBM
    from multinn_b200.params import ParamArena
    arena = ParamArena()
    rbm = RBM(D, H, k=k, arena=arena, name='rbm')
    arena.finalize('cuda', seed=3)
    g = torch.Generator(device='cuda').manual_seed(0)
    res = {}
    default_mode = ops.GIBBS_MODE
    burn = torch.randn(8192, 8192, device='cuda')
    for _ in range(30):                       # leave the idle clocks before the first timed launch
        burn @ burn
    torch.cuda.synchronize()
    # ascending sizes: timed right after the 65 536-row GEMM path the 2 048-row fused case once read 4.1 ms against 0.22 ms
    # standalone (tools/gibbs_dbg.py; caching-allocator traffic inside the timed region), an artefact of the order
    for n_rows, density in ((256, 0.05), (2048, 0.05), (N, 0.05), (N, 0.5)):
        v = (torch.rand(n_rows, D, device='cuda', generator=g) < density).float()
        bh = torch.randn(n_rows, H, device='cuda', generator=g) * 0.3
        bv = torch.randn(n_rows, D, device='cuda', generator=g) * 0.3 - 2.0          # sparse visibles, like piano-rolls
        for mode in ('fused', 'gemm'):
            ops.GIBBS_MODE = mode
            for _ in range(2):
                rbm.sample(v, bh, bv)
            l0 = _lib.lib.mnn_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                p, vk = rbm.sample(v, bh, bv)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            res[f'{mode}_n{n_rows}_d{density}'] = dict(ms=round(ms, 4), launches=(_lib.lib.mnn_launch_count() - l0) // iters,
                                             rows_per_s=round(n_rows / ms * 1e3), mean_vk=round(float(vk.mean()), 5),
                                             mean_p=round(float(p.mean()), 5),
                                             gflops=round(n_rows * k * 4 * D * H / ms / 1e6, 1))
        ops.GIBBS_MODE = default_mode
    del burn
    print(json.dumps({'gibbs_bench': dict(N=N, D=D, H=H, k=k), **res}))
    return res


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 65536)
