"""
Host mirror of the counter-based normal generator fused into the sample kernel
(``philox_normal`` in csrc/common.cuh): Philox4x32-10 keyed by ``seed``, counter
``(n // 2, sequence id)``, two 53-bit uniforms, Box-Muller with both outputs used
(even n -> cos branch, odd n -> sin branch).

``normals(seed, seq, N)`` regenerates exactly the stream the GPU drew for global
sequence ``seq`` (up to libm rounding of log / sincospi), which is how the fused
sampling path is checked against the oracle: the oracle is fed these normals.
"""
import numpy as np

__all__ = ["philox4x32_10", "normals"]

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Counters are uint64 arrays holding 32-bit values."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def normals(seed, seq, N):
    """Standard normal draws n = 0..N-1 of global sequence ``seq`` under ``seed``."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    seq = int(seq) & 0xFFFFFFFFFFFFFFFF
    n = np.arange(N, dtype=np.uint64)
    m = n >> np.uint64(1)
    c = philox4x32_10(m & _MASK, m >> np.uint64(32),
                      np.full(N, seq & 0xFFFFFFFF, dtype=np.uint64),
                      np.full(N, seq >> 32, dtype=np.uint64),
                      seed & 0xFFFFFFFF, seed >> 32)
    a = (c[1] << np.uint64(32)) | c[0]
    b = (c[3] << np.uint64(32)) | c[2]
    two53 = 2.0 ** -53
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 0.5) * two53
    u2 = ((b >> np.uint64(11)).astype(np.float64) + 0.5) * two53
    r = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    # sincospi(2 u2): reduce exactly before multiplying by pi, like the device routine
    x = 2.0 * u2
    k = np.rint(2.0 * x)
    f = (x - 0.5 * k) * np.pi
    s, cs = np.sin(f), np.cos(f)
    q = k.astype(np.int64) & 3
    sn = np.choose(q, [s, cs, -s, -cs])
    co = np.choose(q, [cs, -s, -cs, s])
    del ang
    odd = (n & np.uint64(1)).astype(bool)
    return np.where(odd, r * sn, r * co)
