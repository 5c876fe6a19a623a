"""ORACLE (test infrastructure) -- numpy statement of the counter-based noise source of
the device throughput modes (OX_NOISE_PHILOX / OX_NOISE_PHILOX_HERMITIAN in
include/orphx.h).  This part has no counterpart in the reference (which uses numpy's
global MT19937, maps.py:1577-1578); it is the CPU definition of OUR noise so that the
device path can be checked at 1e-10 end to end in the throughput modes too:
the noise field built here is fed to the restated reference algorithm
(oracle.maps_np.MapGen.map_from_noise).

Philox4x32-10 (Salmon et al. 2011): counter = (pixel_lo, pixel_hi, component, stream),
key = (seed_lo, seed_hi); Box-Muller on two 53-bit uniforms from the four output words.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint64(k0) & _MASK
    k1 = np.uint64(k1) & _MASK
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def normal2(seed, pix, comp, stream):
    """Two independent N(0,1) per (seed, pixel, component, stream)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    pix = np.asarray(pix, dtype=np.uint64)
    x0, x1, x2, x3 = philox4x32_10(pix & _MASK, pix >> _S32, np.uint64(comp), np.uint64(stream),
                                   seed & 0xFFFFFFFF, seed >> 32)
    a = (x0 << _S32) | x1
    b = (x2 << _S32) | x3
    u1 = ((a >> np.uint64(11)) + np.uint64(1)).astype(np.float64) * 2.0 ** -53
    u2 = (b >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    r = np.sqrt(-2.0 * np.log(u1))
    ang = np.pi * (2.0 * u2)
    return r * np.cos(ang), r * np.sin(ang)


def noise_field(seed, ncomp, ny, nx, hermitian=False):
    """Complex white-noise field (ncomp, ny, nx) as the device draws it.
    hermitian=False: independent N(0,1)+iN(0,1) per full-plane pixel (stream 0).
    hermitian=True : R(p') = conj(R(p)); each pair {p,p'} draws once at its smaller linear
    index (stream 1); self-conjugate pixels are real with unit variance, the others have
    variance 1/2 per component."""
    iy, ix = np.mgrid[0:ny, 0:nx]
    p = (iy * nx + ix).astype(np.uint64)
    out = np.empty((ncomp, ny, nx), dtype=np.complex128)
    if not hermitian:
        for c in range(ncomp):
            n1, n2 = normal2(seed, p, c, 0)
            out[c] = n1 + 1j * n2
        return out
    my, mx = (-iy) % ny, (-ix) % nx
    q = (my * nx + mx).astype(np.uint64)
    canon = np.minimum(p, q)
    conj_me = q < p
    selfc = q == p
    for c in range(ncomp):
        n1, n2 = normal2(seed, canon, c, 1)
        re = np.where(selfc, n1, n1 * 0.70710678118654752440)
        im = np.where(selfc, 0.0, np.where(conj_me, -n2, n2) * 0.70710678118654752440)
        out[c] = re + 1j * im
    return out
