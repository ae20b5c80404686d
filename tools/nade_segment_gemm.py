"""Design prototype (NumPy, CPU) for round 2's tensor-core NADE kernels: the teacher-forced NADE log-likelihood AND its
whole backward as FOUR dense GEMMs over "segment rows", plus elementwise work. Not part of the product path.

Why. `NADE.log_prob` (reference common/nade.py:155-229) updates `a += v_i * w_enc[i]` only where the target bit is 1, so
the hidden vector h = sigmoid(a) is piecewise constant over the D dims of a row: a row with K set bits among dims
0..D-2 has K + 1 distinct hidden vectors ("segments"; about 5.2 per row at the 5 % density of piano-rolls, against
D = 84 in the reference's dense loop). Stack the segments of all rows into R ~ 5.2 N segment rows:

    A[r]  = b_enc[n(r)] + sum of the w_enc rows of the first k(r) set bits          (prefix adds, CUDA cores)
    H     = sigmoid(A)                                                   [R, H]
    Lt    = H @ W_dec^T                                                  [R, D]   GEMM 1 (tensor cores)
    l[n,i]= b_dec[n,i] + Lt[r(n, seg(n,i)), i]          each (n,i) reads exactly one entry: segment rows own a
                                                        CONTIGUOUS dim range [lo(r), hi(r)], so the epilogue thread that
                                                        holds TMEM lane r walks its own columns only
    g     = dNLL/dl  (fused sigmoid / BCE-with-eps epilogue), Gt[r,i] = g[n,i] on the row's range, 0 elsewhere
    dW_dec= Gt^T @ H                                                     [D, H]   GEMM 2
    dH    = Gt @ W_dec                                                   [R, H]   GEMM 3
    dA    = dH * H * (1 - H);  S[r] = suffix sum of dA over the later segments of the same row
    d b_enc[n] = S[r(n,0)];    dW_enc = Vt^T @ S                         [D, H]   GEMM 4, Vt[r,j] = 1 iff segment r was
                                                                                   opened by set bit j (binary, exact)
The redundancy is (K+1) ~ 5.2x over the useful decode flops instead of the 84x of the dense triangular contraction:
4 GEMMs x 2 * R * 84 * 256 flops with R = 5.2 * 524 288 * 5 tracks = 2.3 TFLOP per C5 step, about 8 ms with the
2.5-product fp32-accurate scheme of the pair GEMM against 30 ms for today's SIMT kernels.

tests/test_nade_segment_gemm.py checks this formulation against the loop-form oracle (values) and fp64 autograd
(every gradient), including all-zero rows, all-one rows and a set last bit (which opens no segment).
"""
import numpy as np


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def segment_rows(V):
    """V[N,D] in {0,1} -> dict of per-segment-row index arrays:
    n[r] source row, k[r] segment number, lo[r]..hi[r] inclusive dim range, opener[r] = set bit that opened it (-1: k=0).
    Only bits at dims 0..D-2 open a segment (a_{D} is never used)."""
    N, D = V.shape
    n_, k_, lo_, hi_, op_ = [], [], [], [], []
    for n in range(N):
        bits = np.flatnonzero(V[n, :D - 1])
        starts = np.concatenate([[0], bits + 1])
        ends = np.concatenate([bits, [D - 1]])
        for k, (lo, hi) in enumerate(zip(starts, ends)):
            n_.append(n); k_.append(k); lo_.append(lo); hi_.append(hi); op_.append(bits[k - 1] if k else -1)
    return {name: np.asarray(a, np.int64) for name, a in dict(n=n_, k=k_, lo=lo_, hi=hi_, opener=op_).items()}


def forward(V, b_enc, b_dec, w_enc, w_dec, eps=1e-6):
    """Returns (nll[N], cond_p[N,D], cache). Same results as the reference's D-step loop."""
    N, D = V.shape
    seg = segment_rows(V)
    R = seg['n'].size
    A = np.empty((R, w_enc.shape[1]), V.dtype)
    for r in range(R):                                   # prefix adds (a running vector per row in the kernel)
        A[r] = b_enc[seg['n'][r]] if seg['k'][r] == 0 else A[r - 1] + w_enc[seg['opener'][r]]
    H = sigmoid(A)
    Lt = H @ w_dec.T                                     # GEMM 1: [R, D]
    cols = np.arange(D)[None, :]
    own = (cols >= seg['lo'][:, None]) & (cols <= seg['hi'][:, None])      # each (n, i) is owned by exactly one r
    L = np.array(b_dec, dtype=V.dtype, copy=True)
    rr, ii = np.nonzero(own)
    L[seg['n'][rr], ii] += Lt[rr, ii]
    P = sigmoid(L)
    ll = (V * np.log(eps + P) + (1 - V) * np.log(eps + 1 - P)).sum(1)
    return -ll, P, dict(seg=seg, H=H, own=own, P=P)


def backward(V, w_enc, w_dec, cache, dnll, eps=1e-6):
    """Gradient of sum_n dnll[n] * NLL[n] wrt b_enc, b_dec, w_enc, w_dec."""
    seg, H, own, P = cache['seg'], cache['H'], cache['own'], cache['P']
    N, D = V.shape
    # d(-ll)/dl = -(v / (eps + p) - (1 - v) / (eps + 1 - p)) * p * (1 - p)
    g = -(V / (eps + P) - (1 - V) / (eps + 1 - P)) * P * (1 - P) * dnll[:, None]
    Gt = np.where(own, g[seg['n']], 0.0)                 # [R, D]
    dW_dec = Gt.T @ H                                    # GEMM 2
    dH = Gt @ w_dec                                      # GEMM 3
    dA = dH * H * (1 - H)
    S = dA.copy()
    for r in range(S.shape[0] - 2, -1, -1):              # suffix sums within a source row
        if seg['n'][r + 1] == seg['n'][r]:
            S[r] += S[r + 1]
    d_b_enc = np.zeros((N, H.shape[1]), V.dtype)
    first = seg['k'] == 0
    d_b_enc[seg['n'][first]] = S[first]
    Vt = np.zeros((S.shape[0], D), V.dtype)
    opened = ~first
    Vt[np.flatnonzero(opened), seg['opener'][opened]] = 1.0
    dW_enc = Vt.T @ S                                    # GEMM 4
    return dict(b_enc=d_b_enc, b_dec=g, w_enc=dW_enc, w_dec=dW_dec)


def pack_tiles(V, tile_rows=128):
    """Tiling rule for the kernels: consecutive source rows are packed into tiles of `tile_rows` segment rows (= TMEM
    lanes of one accumulator) and a source row never straddles two tiles, so the suffix sums of the backward and the
    per-row NLL reduction stay inside one CTA. Returns (first source row of every tile + end sentinel, fill ratio)."""
    N, D = V.shape
    per_row = V[:, :D - 1].sum(1).astype(np.int64) + 1
    assert per_row.max() <= tile_rows, "a row's segments must fit one tile (D <= tile_rows)"
    starts, used, total = [0], 0, 0
    for n in range(N):
        if used + per_row[n] > tile_rows:
            starts.append(n)
            total += used
            used = 0
        used += per_row[n]
    total += used
    starts.append(N)
    return np.asarray(starts), total / ((len(starts) - 1) * tile_rows)


def forward_backward_tiled(V, b_enc, b_dec, w_enc, w_dec, dnll, tile_rows=128, kblock=32, flush_every=16, eps=1e-6):
    """The same computation in the ORDER the planned kernels do it, tile by tile (one tile = `tile_rows` segment rows =
    the TMEM lanes of one accumulator, whole source rows only):
      producer   per k-block of `kblock` hidden units: walk each source row's set bits with a `kblock`-wide slice of `a`
                 and emit one H row per segment (what the producer warps write into the A-operand stage);
      MMA 1      Lt += H[:, kb] @ W_dec[:, kb]^T over the k-blocks                      [tile_rows, D]
      epilogue   lane r owns dims lo(r)..hi(r): logits, probabilities, NLL partials, g; per-source-row NLL = sum of its
                 lanes' partials (stays inside the tile because rows never straddle);
      backward   Gt (masked g) -> MMA 3: dH = Gt @ W_dec; dA = dH*H*(1-H); suffix sums inside the tile -> d b_enc, S;
                 MMA 2 / 4: dW_dec += Gt^T @ H, dW_enc += Vt^T @ S into PERSISTENT accumulators that are flushed (added
                 into the global result) every `flush_every` tiles - the TMEM-resident accumulators of the CTA.
    Returns (nll, cond_p, grads) equal to forward() / backward()."""
    N, D = V.shape
    Hdim = w_enc.shape[1]
    assert Hdim % kblock == 0
    starts, _ = pack_tiles(V, tile_rows)
    nll = np.zeros(N, V.dtype)
    P = np.zeros((N, D), V.dtype)
    g_b_enc, g_b_dec = np.zeros((N, Hdim), V.dtype), np.zeros((N, D), V.dtype)
    g_w_enc, g_w_dec = np.zeros_like(w_enc), np.zeros_like(w_dec)
    acc_enc, acc_dec, pending = np.zeros_like(w_enc), np.zeros_like(w_dec), 0
    cols = np.arange(D)[None, :]
    for a, b in zip(starts[:-1], starts[1:]):
        rows = np.arange(a, b)
        bits = [np.flatnonzero(V[n, :D - 1]) for n in rows]
        r_of, lo, hi, opener, first_lane = [], [], [], [], {}
        for j, n in enumerate(rows):                                  # lane table of the tile (built once per tile)
            first_lane[n] = len(r_of)
            for k in range(len(bits[j]) + 1):
                r_of.append(n)
                lo.append(0 if k == 0 else bits[j][k - 1] + 1)
                hi.append(bits[j][k] if k < len(bits[j]) else D - 1)
                opener.append(-1 if k == 0 else bits[j][k - 1])
        R = len(r_of)
        assert R <= tile_rows
        r_of, lo, hi, opener = (np.asarray(x) for x in (r_of, lo, hi, opener))
        # ---- producer + MMA 1, k-block by k-block
        Htile = np.zeros((R, Hdim), V.dtype)
        Lt = np.zeros((R, D), V.dtype)
        for kb in range(0, Hdim, kblock):
            sl = slice(kb, kb + kblock)
            for j, n in enumerate(rows):
                avec = b_enc[n, sl].copy()
                r = first_lane[n]
                Htile[r, sl] = sigmoid(avec)
                for bit in bits[j]:
                    avec = avec + w_enc[bit, sl]
                    r += 1
                    Htile[r, sl] = sigmoid(avec)
            Lt += Htile[:, sl] @ w_dec[:, sl].T
        # ---- epilogue: lane r walks its own dims only
        own = (cols >= lo[:, None]) & (cols <= hi[:, None])
        L = b_dec[r_of] + Lt
        Pt = sigmoid(L)
        Vr = V[r_of]
        part = -np.where(own, Vr * np.log(eps + Pt) + (1 - Vr) * np.log(eps + 1 - Pt), 0.0).sum(1)
        Gt = np.where(own, -(Vr / (eps + Pt) - (1 - Vr) / (eps + 1 - Pt)) * Pt * (1 - Pt) * dnll[r_of][:, None], 0.0)
        for n in rows:                                                # the lane with k == 0 sums its row's partials
            lanes = np.flatnonzero(r_of == n)
            nll[n] = part[lanes].sum()
            P[n] = (Pt[lanes] * own[lanes]).sum(0)                    # each dim is owned by exactly one lane
            g_b_dec[n] = Gt[lanes].sum(0)
        # ---- backward
        dH = Gt @ w_dec                                               # MMA 3
        S = dH * Htile * (1 - Htile)
        for r in range(R - 2, -1, -1):
            if r_of[r + 1] == r_of[r]:
                S[r] += S[r + 1]
        for n in rows:
            g_b_enc[n] = S[first_lane[n]]
        Vt = np.zeros((R, D), V.dtype)
        opened = opener >= 0
        Vt[np.flatnonzero(opened), opener[opened]] = 1.0
        acc_dec += Gt.T @ Htile                                       # MMA 2 (persistent accumulator)
        acc_enc += Vt.T @ S                                           # MMA 4
        pending += 1
        if pending == flush_every:                                    # cp.reduce.async.bulk add into global memory
            g_w_dec += acc_dec
            g_w_enc += acc_enc
            acc_dec[:], acc_enc[:], pending = 0.0, 0.0, 0
    g_w_dec += acc_dec
    g_w_enc += acc_enc
    return nll, P, dict(b_enc=g_b_enc, b_dec=g_b_dec, w_enc=g_w_enc, w_dec=g_w_dec)


def flops(N, D, H, K_mean, tracks=1):
    R = (K_mean + 1) * N * tracks
    return dict(segment_rows=R, gemm_flops=4 * 2 * R * D * H, dense_triangular_flops=3 * tracks * N * D * (D - 1) * H)


if __name__ == '__main__':
    f = flops(524288, 84, 256, 0.05 * 83, tracks=5)
    print({k: f'{v:.3e}' for k, v in f.items()})
    V = (np.random.default_rng(23).random((20000, 84)) < 0.05).astype(np.float32)
    print('tile fill at 5 % density:', round(pack_tiles(V)[1], 4))
