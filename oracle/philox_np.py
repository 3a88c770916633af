"""TEST INFRASTRUCTURE -- numpy restatement of the Philox4x32-10 dropout masks the CUDA kernels draw
(pixelrec_b200/csrc/common.cuh `Philox`, ln.cu `drop4`, attn.cu `attn_keep`), so parity tests can run
WITH dropout: the oracle multiplies by exactly the mask the kernel used.  The reference's own dropout
(torch RNG) cannot be reproduced bit for bit; what is checked is the reference's
arithmetic given the same mask.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint64(0x9E3779B9)
W1 = np.uint64(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, stream, seed):
    """ctr: uint64 array; returns uint32 array [..., 4]."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    c0 = ctr & MASK
    c1 = (ctr >> np.uint64(32)) & MASK
    c2 = np.full_like(c0, np.uint64(stream))
    c3 = np.zeros_like(c0)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = (p0 >> np.uint64(32)) & MASK, p0 & MASK
        hi1, lo1 = (p1 >> np.uint64(32)) & MASK, p1 & MASK
        n0 = hi1 ^ c1 ^ k0
        n2 = hi0 ^ c3 ^ k1
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return np.stack([c0, c1, c2, c3], -1).astype(np.uint32)


def drop_threshold(p):
    t = float(np.float32(p)) * 4294967296.0
    return np.uint32(min(max(t, 0.0), 4294967295.0))


def rowwise_keep_scale(rows, D, p, seed, stream, dtype=np.float32):
    """Multiplicative mask [rows, D] of pr_add_ln_*: element (r, 4c+e) <- Philox(ctr=r*D/4+c)[e]."""
    if p <= 0:
        return np.ones((rows, D), dtype=dtype)
    D4 = D // 4
    ctr = (np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(D4) + np.arange(D4, dtype=np.uint64)[None, :])
    r = philox4x32_10(ctr, stream, seed).reshape(rows, D)
    keep = r >= drop_threshold(p)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(keep, inv, np.float32(0)).astype(dtype)


def attn_keep_scale(B, h, L, p, seed, stream, dtype=np.float32):
    """Mask [B,h,L,L] of pr_sasrec_attn_*: entry (item, i, j) <- Philox(ctr=((item*L+i)*8 + j%8)*2 + (j//8)//4)[(j//8)%4]."""
    if p <= 0:
        return np.ones((B, h, L, L), dtype=dtype)
    item = np.arange(B * h, dtype=np.uint64)[:, None, None]
    i = np.arange(L, dtype=np.uint64)[None, :, None]
    j = np.arange(L, dtype=np.uint64)[None, None, :]
    jj = j // np.uint64(8)
    ctr = ((item * np.uint64(L) + i) * np.uint64(8) + (j % np.uint64(8))) * np.uint64(2) + jj // np.uint64(4)
    r = philox4x32_10(ctr, stream, seed)                       # [B*h, L, L, 4]
    comp = np.broadcast_to((jj % np.uint64(4)).astype(np.int64), ctr.shape)
    rr = np.take_along_axis(r, comp[..., None], -1)[..., 0]
    keep = rr >= drop_threshold(p)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(keep, inv, np.float32(0)).astype(dtype).reshape(B, h, L, L)
