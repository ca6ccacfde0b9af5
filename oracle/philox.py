"""NumPy restatement of libauxssm's counter-based draws (oracle only; test infrastructure).

Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011): counter words (block index lo, block index hi, call lo, call hi),
key words (seed lo, seed hi).  Draw call number `call` fills an array of n entries:
  fp32: 4 entries per block; uniform = ((word >> 8) + 0.5) / 2^24; normals = Box-Muller on (w0, w1) and (w2, w3):
        sqrt(-2 log a) * (cos 2 pi b, sin 2 pi b)
  fp64: 2 entries per block; uniform = ((w0 << 32 | w1) >> 11) + 0.5) / 2^53 and likewise (w2, w3); one Box-Muller
        pair per block.
Uniforms are bit-exact with the kernels; normals agree to a few ulp (log / sincospi rounding).
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(blocks, call, seed):
    """blocks: uint64 array of block indices -> uint32 array [len, 4]."""
    blocks = np.asarray(blocks, dtype=np.uint64)
    c0, c1 = blocks & _MASK, blocks >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(call & 0xFFFFFFFF))
    c3 = np.full_like(c0, np.uint64((call >> 32) & 0xFFFFFFFF))
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK, p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _u01(words, dtype):
    if dtype == np.float32:
        return ((words >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    w = words.astype(np.uint64)
    m = ((w[..., 0::2] << np.uint64(32)) | w[..., 1::2]) >> np.uint64(11)
    return (m.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def fill(kind, shape, dtype, seed, call):
    dtype = np.dtype(dtype).type
    n = int(np.prod(shape))
    per = 4 if dtype == np.float32 else 2
    nblk = (n + per - 1) // per
    w = philox4x32_10(np.arange(nblk, dtype=np.uint64), call, seed)
    u = _u01(w, dtype)                              # [nblk, per]
    if kind == "uniform":
        out = u
    else:
        a, b = u[:, 0::2], u[:, 1::2]
        r = np.sqrt(dtype(-2) * np.log(a))
        ang = dtype(2 * np.pi) * b
        out = np.stack([r * np.cos(ang), r * np.sin(ang)], axis=-1).reshape(nblk, per)
    return out.reshape(-1)[:n].reshape(shape).astype(dtype)


class PhiloxDraws:
    """CPU twin of aux_ssm_samplers_b200.draws.PhiloxDraws (same call-counter convention)."""

    def __init__(self, seed):
        self.seed, self.call = int(seed) & 0xFFFFFFFFFFFFFFFF, 0

    def normal(self, shape, dtype):
        out = fill("normal", shape, dtype, self.seed, self.call)
        self.call += 1
        return out

    def uniform(self, shape, dtype):
        out = fill("uniform", shape, dtype, self.seed, self.call)
        self.call += 1
        return out
