"""Times the GEMM shapes of one Composer C5 training step on the tcgen05 3xTF32 kernel vs the CUDA-core kernel.
Run on a B200: python tools/gemm_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200 import ops  # noqa: E402

N = 524288
SHAPES = [  # name, M, N, K, transA, transB, a_exact
    ('L1 xproj fwd', N, 2048, 420, 0, 0, 1), ('L2 xproj fwd', N, 1024, 512, 0, 0, 0), ('Dense fwd', N, 1700, 256, 0, 0, 0),
    ('dout = dfc K^T', N, 256, 1700, 0, 1, 0), ('d_in2 = dG2 W2x^T', N, 512, 1024, 0, 1, 0),
    ('dW1x', 420, 2048, N, 1, 0, 1), ('dW1h', 512, 2048, N, 1, 0, 0), ('dW2x', 512, 1024, N, 1, 0, 0),
    ('dW2h', 256, 1024, N, 1, 0, 0), ('dK', 256, 1700, N, 1, 0, 0),
    ('L1 recur fwd (1 step)', 2048, 2048, 512, 0, 0, 0), ('L2 recur fwd (1 step)', 2048, 1024, 256, 0, 0, 0),
    ('L1 recur bwd (1 step)', 2048, 512, 2048, 0, 1, 0), ('L2 recur bwd (1 step)', 2048, 256, 1024, 0, 1, 0),
    ('L1 recur fwd B=256', 256, 2048, 512, 0, 0, 0),
]


def main():
    only = sys.argv[1:] or ['tc', 'f32']
    pre = os.environ.get('GB_PRE') == '1'      # pair split with the weight operand pre-split (training-step path)
    if os.environ.get('GB_PAIR') == '1' or pre:
        ops.set_gemm_split('pair')
        ops._gemm_split = 'pair'
    pick = os.environ.get('GB_ONLY')
    for name, M, Nn, K, ta, tb, ex in SHAPES:
        if pick and not any(t in name for t in pick.split(',')):
            continue
        A = torch.randn((K, M) if ta else (M, K), device='cuda')
        B = torch.randn((Nn, K) if tb else (K, Nn), device='cuda')
        C = torch.empty(M, Nn, device='cuda')
        line = f'{name:24s} M={M:7d} N={Nn:5d} K={K:7d}'
        for mode in only:
            if mode == 'f32' and M * Nn * K > 3e14:
                continue
            reps = 3 if M * Nn * K > 1e11 else 20
            for _ in range(2):
                ops.gemm(A, B, C, transA=bool(ta), transB=bool(tb), a_exact=bool(ex), mode=mode, b_weight=pre and not ta)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ops.gemm(A, B, C, transA=bool(ta), transB=bool(tb), a_exact=bool(ex), mode=mode, b_weight=pre and not ta)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            line += f' | {mode}: {ms:8.3f} ms {2.0 * M * Nn * K / ms / 1e9:7.1f} TFLOP/s'
        print(line, flush=True)
        del A, B, C


if __name__ == '__main__':
    main()
