import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinn_b200 import ops

def run(M, N, K, ta, tb, a_exact=False):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    if a_exact: A = (A > 0.8).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    ref = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
    C = torch.zeros(M, N, device='cuda')
    ops.gemm(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), C, transA=bool(ta), transB=bool(tb), a_exact=a_exact, mode='tc')
    torch.cuda.synchronize()
    got = C.cpu().numpy().astype(np.float64)
    err = np.abs(got - ref)
    bad = err > 1e-3 * np.sqrt(K)
    msg = f'M={M} N={N} K={K} ta={ta} tb={tb}: maxerr/sqrtK={err.max()/np.sqrt(K):.3e} bad={bad.mean():.4f}'
    if bad.any():
        rows = np.where(bad.any(1))[0]; cols = np.where(bad.any(0))[0]
        msg += f' badrows[{rows.min()}..{rows.max()}] n={len(rows)} badcols[{cols.min()}..{cols.max()}] n={len(cols)}'
        r, c = rows[0], cols[0]
        msg += f' got[{r},{c}]={got[r,c]:.4f} ref={ref[r,c]:.4f}'
    print(msg, flush=True)

for ta, tb in [(0,1),(0,0),(1,1),(1,0)]:
    for (M,N,K) in [(128,128,32),(128,128,8),(128,64,32),(128,256,64),(300,200,100),(77,340,420),(1000,1700,256),(2048,2048,512)]:
        run(M,N,K,ta,tb)
run(512,2048,420,0,0,True)
run(420,2048,16384,1,0)
