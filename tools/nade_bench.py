"""NADE forward kernels at the C5 row count: SIMT vs tcgen05 (time, agreement), by target density.
  python tools/nade_bench.py [N] [density ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200 import ops  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
    dens = [float(a) for a in sys.argv[2:]] or [0.05]
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
        res = {}
        for mode in ('simt', 'tc'):
            ops.set_nade_mode(mode)
            nll = torch.empty(M, N, device='cuda')
            dfc = torch.zeros_like(fc)
            for _ in range(2):
                ops.nade_logprob_fwd(bits, fc, 0, M * H, we, wd, nll, dfc=dfc, gscale=1.0 / (N * M))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.nade_logprob_fwd(bits, fc, 0, M * H, we, wd, nll, dfc=dfc, gscale=1.0 / (N * M))
            e1.record()
            torch.cuda.synchronize()
            res[mode] = (e0.elapsed_time(e1) / 5, nll.clone(), dfc.clone())
        ops.set_nade_mode('simt')
        (ts, ns, ds), (tt, nt, dt) = res['simt'], res['tc']
        rel = float(((ns - nt).abs() / ns.abs()).max())
        drel = float((ds - dt).abs().max() / ds.abs().max())
        print(f'N={N} density={density}: simt {ts:.3f} ms, tc {tt:.3f} ms, max rel dNLL {rel:.2e}, max d(dl)/max {drel:.2e}',
              flush=True)


if __name__ == '__main__':
    main()
