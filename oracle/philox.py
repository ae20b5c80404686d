"""TEST INFRASTRUCTURE ONLY. NumPy Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11)
and the counter layouts of the in-kernel generators (csrc/common.cuh::philox4x32_10, csrc/rbm.cu), so that tests can
reproduce on the CPU exactly the uniforms a kernel draws in Philox mode and check sampled outputs bit for bit.
Pinned by the Random123 known-answer vectors in tests/test_oracle_kat.py."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr[..., 4], key[..., 2] uint32 (broadcastable) -> [..., 4] uint32."""
    ctr = np.asarray(ctr, np.uint32)
    key = np.asarray(key, np.uint32)
    c = [np.array(ctr[..., i], np.uint32) for i in range(4)]
    k0, k1 = np.array(key[..., 0], np.uint32), np.array(key[..., 1], np.uint32)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0, k1 = (k0 + W0).astype(np.uint32), (k1 + W1).astype(np.uint32)
    return np.stack(np.broadcast_arrays(*c), axis=-1)


def u01(bits):
    """common.cuh::u01: the top 24 bits as a float32 in [0, 1)."""
    return ((np.asarray(bits, np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0))


def _split64(x):
    x = np.asarray(x, np.uint64)
    return (x & MASK).astype(np.uint32), (x >> np.uint64(32)).astype(np.uint32)


def _global_rows(n, row_map):
    """common.cuh::global_row for local rows 0..n-1; row_map = (rows_local, rows_global, row_base) or None."""
    r = np.arange(n, dtype=np.uint64)
    if not row_map or not row_map[0]:
        return r
    rl, rg, rb = (np.uint64(v) for v in row_map)
    return (r // rl) * rg + rb + r % rl


def half_step_uniforms(seed, offset, N, C, row_map=None):
    """csrc/rbm.cu::bias_sigmoid_sample_kernel in Philox mode: element (r, c) draws word 0 of
    Philox(ctr = (offset + global_row(r)*C + c as 64 bits, 0, 0), key = seed as 64 bits). Returns float32 [N, C]."""
    g = _global_rows(N, row_map)
    e = np.uint64(offset) + (g[:, None] * np.uint64(C) + np.arange(C, dtype=np.uint64)[None, :]).reshape(-1)
    lo, hi = _split64(e)
    klo, khi = _split64(np.uint64(seed))
    z = np.zeros_like(lo)
    out = philox4x32_10(np.stack([lo, hi, z, z], -1), np.stack([klo, khi], -1)[None, :])
    return u01(out[:, 0]).reshape(N, C)


def gibbs_chain_uniforms(seed, offset, N, D, H, k, row_map=None):
    """csrc/rbm.cu::rbm_gibbs_kernel in Philox mode: row r, half-step hs = 2*s (hidden) or 2*s + 1 (visible), column c
    draws word c % 4 of Philox(ctr = (offset + r as 64 bits, hs, c // 4), key = seed). Returns (uh[k,N,H], uv[k,N,D])."""
    klo, khi = _split64(np.uint64(seed))
    key = np.stack([klo, khi], -1)

    def draw(hs, C):
        lo, hi = _split64(np.uint64(offset) + _global_rows(N, row_map))
        g = np.arange(C // 4, dtype=np.uint32)
        ctr = np.stack(np.broadcast_arrays(lo[:, None], hi[:, None], np.uint32(hs), g[None, :]), -1)      # [N, C/4, 4]
        return u01(philox4x32_10(ctr, key)).reshape(N, C)

    uh = np.stack([draw(2 * s, H) for s in range(k)])
    uv = np.stack([draw(2 * s + 1, D) for s in range(k)])
    return uh, uv


def nade_sample_uniforms(seed, offset, M, N, D, row_map=None):
    """csrc/nade.cu::nade_sample_kernel (and csrc/sample.cu) in Philox mode: (track m, row n, dim i) draws word 0 of
    Philox(ctr = ((global_row(n)*M + m)*D + i as 64 bits, offset as 64 bits), key = seed). Returns float32 [M, N, D]
    (`offset` is the generated step index in RnnEstimator.generate)."""
    g = _global_rows(N, row_map)
    idx = ((g[None, :, None] * np.uint64(M) + np.arange(M, dtype=np.uint64)[:, None, None]) * np.uint64(D)
           + np.arange(D, dtype=np.uint64)[None, None, :]).reshape(-1)
    lo, hi = _split64(idx)
    olo, ohi = _split64(np.uint64(offset))
    klo, khi = _split64(np.uint64(seed))
    ctr = np.stack(np.broadcast_arrays(lo, hi, olo, ohi), -1)
    return u01(philox4x32_10(ctr, np.stack([klo, khi], -1)[None, :])[:, 0]).reshape(M, N, D)


def dropout_uniforms(seed, T, B, R, row_map=None, t_base=0):
    """csrc/common.cuh::dropout_bits4 -- the dropout noise of every LSTM kernel (SIMT cell, tcgen05 sequence kernels, any
    chunking): units [4q, 4q+4) of batch row b at step t are the 4 words of
    Philox(ctr = (global_row(b) * R/4 + q as 64 bits, t_base + t, 0), key = seed). Returns float32 [T, B, R]."""
    g = _global_rows(B, row_map)
    e = (g[:, None] * np.uint64(R // 4) + np.arange(R // 4, dtype=np.uint64)[None, :]).reshape(-1)         # [B*R/4]
    lo, hi = _split64(e)
    klo, khi = _split64(np.uint64(seed))
    key = np.stack([klo, khi], -1)[None, :]
    out = np.empty((T, B, R), np.float32)
    for t in range(T):
        ctr = np.stack(np.broadcast_arrays(lo, hi, np.uint32(t_base + t), np.uint32(0)), -1)
        out[t] = u01(philox4x32_10(ctr, key)).reshape(B, R)
    return out
