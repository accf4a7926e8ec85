"""Oracle for the in-kernel noise of psx_dps_post_philox (numpy; TEST INFRASTRUCTURE).

Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 v1.14 known-answer
vectors below) -- the counter-based generator family torch's CUDA ``randn`` uses (the reference's
``torch.randn_like``, bridge_kernels.py:59); the stream layout is this library's own, so results are NOT
bit-comparable with torch ("production mode ... not bit-comparable"):

  key     = (seed_lo, seed_hi)
  counter = (g_lo, g_hi, step_lo, step_hi),  g = flat element index // 4  (element = l * n + i)
  the four outputs feed elements 4g .. 4g+3:  u = (x + 0.5) * 2^-32 computed in fp32 as x * 2^-32 + 2^-33,
  (z0, z1) = Box-Muller(u0, u1), (z2, z3) = Box-Muller(u2, u3) with r = sqrt(-2 ln u_a), theta = 2 pi u_b,
  z_a = r cos(theta), z_b = r sin(theta).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter (..., 4) uint32, key (2,) uint32 -> (..., 4) uint32."""
    c = [counter[..., i].astype(np.uint32) for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for r in range(10):
            if r:
                k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
                k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
            p0 = c[0].astype(np.uint64) * M0
            p1 = c[2].astype(np.uint64) * M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
    return np.stack(c, axis=-1)


def uniforms(bits: np.ndarray) -> np.ndarray:
    """uint32 -> fp32 in (0, 1]: x * 2^-32 + 2^-33, each op rounded to fp32 (as the kernel's FFMA-free sequence)."""
    x = bits.astype(np.float32)                      # round-to-nearest conversion, as cvt.rn.f32.u32
    return (x * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float32)


def normals(numel: int, seed: int, step: int) -> np.ndarray:
    """The N(0,1) field the kernel adds at ``step`` for a state of ``numel`` elements (fp32)."""
    groups = (numel + 3) // 4
    g = np.arange(groups, dtype=np.uint64)
    ctr = np.stack([(g & MASK).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                    np.full(groups, step & 0xFFFFFFFF, dtype=np.uint32),
                    np.full(groups, (step >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    u = uniforms(philox4x32_10(ctr, key)).astype(np.float64)
    out = np.empty((groups, 4), dtype=np.float64)
    for a, b in ((0, 1), (2, 3)):
        r = np.sqrt(-2.0 * np.log(u[:, a]))
        th = 2.0 * np.pi * u[:, b]
        out[:, a], out[:, b] = r * np.cos(th), r * np.sin(th)
    return out.reshape(-1)[:numel].astype(np.float32)


# Random123 kat_vectors, philox4x32 10 rounds: (counter, key, expected)
KAT = [
    ((0x00000000,) * 4, (0x00000000,) * 2, (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]
