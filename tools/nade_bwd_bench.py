"""NADE backward kernel alone at the C5 row count, by target density (event-timed; the ncu source capture of
profiles/r2_nade_bwd_* runs this with one iteration).
  python tools/nade_bwd_bench.py [N] [iters] [density ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200 import ops  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dens = [float(a) for a in sys.argv[3:]] or [0.05]
    M, D, H = 5, 84, 256
    g = torch.Generator(device='cuda').manual_seed(0)
    fc = torch.randn(N, M * (H + D), device='cuda', generator=g)
    fc[:, M * H:] -= 2.0
    we = torch.randn(M, D, H, device='cuda', generator=g) / D ** 0.5
    wd = torch.randn(M, D, H, device='cuda', generator=g) / D ** 0.5
    for density in dens:
        x = (torch.rand(M, N, D, device='cuda', generator=g) < density).float()
        bits = torch.empty(M, N, 4, dtype=torch.int32, device='cuda')
        for m in range(M):
            ops.pack_rows(x[m], bits[m], D)
        nll = torch.empty(M, N, device='cuda')
        dfc = torch.zeros_like(fc)
        dwe, dwd = torch.zeros_like(we), torch.zeros_like(wd)
        ops.nade_logprob_fwd(bits, fc, 0, M * H, we, wd, nll, dfc=dfc, gscale=1.0 / (N * M))
        for _ in range(min(iters, 2)):
            ops.nade_logprob_bwd(bits, fc, 0, M * H, we, wd, dfc, dwe, dwd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.nade_logprob_bwd(bits, fc, 0, M * H, we, wd, dfc, dwe, dwd)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 4.0 * N * M * D * H
        print(f'N={N} density={density}: nade_bwd {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s useful '
              f'(checksum dW_enc {float(dwe.double().abs().sum()):.6e} dW_dec {float(dwd.double().abs().sum()):.6e} '
              f'd b_enc {float(dfc[:, :M * H].double().abs().sum()):.6e})', flush=True)


if __name__ == '__main__':
    main()
