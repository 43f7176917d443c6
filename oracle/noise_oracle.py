"""ORACLE (test infrastructure, never shipped, never on the product path).

numpy restatement of `egr_noise_fill` (csrc/core.cu): the FlashSR sampler's start noise x_T.  Upstream draws it with
torch.randn inside FlashSR.forward (reference call site egregora_audio_super_resolution.py:366-369 — unseeded, so no
reference vector exists: **parity unpinned** at that boundary).  The build defines x_T as a pure
function of (seed, global chunk-channel row, element): Philox4x32-10 (Salmon et al., SC'11 — known-answer vectors of
the Random123 distribution are checked in tests/test_noise.py) -> u = ((bits >> 8) + 1) * 2^-24 -> Box-Muller.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c, k):
    """c: uint32 array [..., 4], k: (k0, k1) python ints -> uint32 [..., 4]."""
    c0, c1, c2, c3 = (c[..., i].astype(np.uint64) for i in range(4))
    k0, k1 = int(k[0]) & 0xFFFFFFFF, int(k[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], -1).astype(np.uint32)


def noise_rows(seed: int, row0: int, n_rows: int, row_elems: int) -> np.ndarray:
    """[n_rows, row_elems] float32 standard normal, the values egr_noise_fill writes (float64 math, rounded once)."""
    quads = (row_elems + 3) // 4
    q = np.arange(quads, dtype=np.uint64)
    out = np.empty((n_rows, quads * 4), np.float32)
    for r in range(n_rows):
        row = np.uint64(row0 + r)
        c = np.stack([q & MASK, q >> np.uint64(32), np.full_like(q, row & MASK), np.full_like(q, row >> np.uint64(32))], -1)
        x = philox4x32_10(c.astype(np.uint32), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
        u = ((x >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
        rad = np.sqrt(-2.0 * np.log(u[:, 0::2]))
        ang = 2.0 * np.pi * u[:, 1::2]
        z = np.empty((quads, 4))
        z[:, 0::2] = rad * np.cos(ang)
        z[:, 1::2] = rad * np.sin(ang)
        out[r] = z.reshape(-1).astype(np.float32)
    return out[:, :row_elems]
