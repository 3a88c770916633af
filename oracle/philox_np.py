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


def philox4x32_10(ctr, stream, seed, c3=0):
    """ctr: uint64 array (counter words 0,1); stream = counter word 2; c3 = counter word 3 (the kernels always use 0; the
    argument exists so the Random123 known-answer vectors can be checked).  Returns uint32 array [..., 4]."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    c0 = ctr & MASK
    c1 = (ctr >> np.uint64(32)) & MASK
    c2 = np.full_like(c0, np.uint64(stream))
    c3 = np.full_like(c0, np.uint64(c3))
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
    """16-bit threshold: keep iff field >= round(p * 2^16)  (common.cuh drop_threshold)."""
    t = float(np.float32(p)) * 65536.0 + 0.5
    return np.uint32(min(max(t, 0.0), 65535.0))


def _fields8(r):
    """[..., 4] uint32 words -> [..., 8] 16-bit fields in keep_bits8 order: low halves of w0..w3, then high halves."""
    return np.concatenate([r & np.uint32(0xFFFF), r >> np.uint32(16)], -1)


def rowwise_keep_scale(rows, D, p, seed, stream, dtype=np.float32):
    """Multiplicative mask [rows, D] of pr_add_ln_*: float4 column c = lane + 32*j (lane = c % 32, j = c // 32),
    element e: field (4*(j&1) + e) of Philox(ctr = row*D/4 + lane + 64*(j>>1))   (ln.cu row_keep_bits)."""
    if p <= 0:
        return np.ones((rows, D), dtype=dtype)
    D4 = D // 4
    c = np.arange(D4, dtype=np.uint64)
    lane, j = c % np.uint64(32), c // np.uint64(32)
    ctr = np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(D4) + (lane + np.uint64(64) * (j >> np.uint64(1)))[None, :]
    f = _fields8(philox4x32_10(ctr, stream, seed))                     # [rows, D4, 8]
    half = (j & np.uint64(1)).astype(np.int64)                          # [D4]
    sel = (4 * half[:, None] + np.arange(4)[None, :])                   # [D4, 4]
    fld = np.take_along_axis(f, np.broadcast_to(sel[None], (rows, D4, 4)), -1).reshape(rows, D)
    keep = fld >= drop_threshold(p)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(keep, inv, np.float32(0)).astype(dtype)


def attn_keep_scale(B, h, L, p, seed, stream, dtype=np.float32):
    """Mask [B,h,L,L] of pr_sasrec_attn_*: entry (item, i, j) <- field (j // 8) of Philox(ctr = (item*L + i)*8 + j % 8)."""
    if p <= 0:
        return np.ones((B, h, L, L), dtype=dtype)
    item = np.arange(B * h, dtype=np.uint64)[:, None, None]
    i = np.arange(L, dtype=np.uint64)[None, :, None]
    j = np.arange(L, dtype=np.uint64)[None, None, :]
    ctr = (item * np.uint64(L) + i) * np.uint64(8) + (j % np.uint64(8))
    f = _fields8(philox4x32_10(ctr, stream, seed))                     # [B*h, L, L, 8]
    jj = np.broadcast_to((j // np.uint64(8)).astype(np.int64), ctr.shape)
    fld = np.take_along_axis(f, jj[..., None], -1)[..., 0]
    keep = fld >= drop_threshold(p)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(keep, inv, np.float32(0)).astype(dtype).reshape(B, h, L, L)
