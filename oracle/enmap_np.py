"""ORACLE (test infrastructure) -- numpy restatement of the pixell.enmap calls the
orphics hot path relies on.

PARITY UNPINNED: pixell is a third-party dependency of the reference that is
neither vendored nor pinned (absent from /root/reference/requirements.txt) and is
not installed here.  Each function below restates pixell's published behaviour
and names the reference call site that depends on it.

Call sites in the reference (all under /root/reference/orphics/):
  enmap.geometry        maps.py:1490
  enmap.area            maps.py:1567, 1592, 1605
  enmap.lmap/modlmap    maps.py:1374, 1607, 1938 ; laxes maps.py:1939
  enmap.fft / ifft      maps.py:1578, 1585, 1613, 1636 ; pixell.fft maps.py:1633, 1923
  enmap.rand_gauss(_harm) maps.py:1578
  enmap.map_mul         maps.py:1579, 1615
  enmap.spec2flat       maps.py:1573, 1592
  enmap.multi_pow       maps.py:1571
  enmap.queb_rotmat     maps.py:1607
  enmap.harm2map        maps.py:1587
"""
import numpy as np
import scipy.fft
import scipy.ndimage

degree = np.pi / 180.0
arcmin = degree / 60.0


class FlatWCS:
    """Minimal stand-in for the astropy WCS pixell carries: a plate-carree (CAR)
    projection described by cdelt/crval/crpix in degrees, [x(ra), y(dec)] order,
    1-based crpix as in FITS."""

    def __init__(self, cdelt, crval, crpix, ctype=("RA---CAR", "DEC--CAR")):
        self.cdelt = np.array(cdelt, dtype=np.float64)
        self.crval = np.array(crval, dtype=np.float64)
        self.crpix = np.array(crpix, dtype=np.float64)
        self.ctype = tuple(ctype)

    @property
    def wcs(self):  # so that wcs.wcs.cdelt works as with astropy
        return self

    def __repr__(self):
        return "car:{cdelt:[%.4g,%.4g],crval:[%.4g,%.4g],crpix:[%.2f,%.2f]}" % (
            self.cdelt[0], self.cdelt[1], self.crval[0], self.crval[1], self.crpix[0], self.crpix[1])


class ndmap(np.ndarray):
    """numpy array carrying a wcs, like pixell.enmap.ndmap."""

    def __new__(cls, arr, wcs=None):
        obj = np.asarray(arr).view(cls)
        obj.wcs = wcs
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.wcs = getattr(obj, "wcs", None)


def enmap(arr, wcs=None):
    return ndmap(np.array(arr), wcs)


def samewcs(arr, *maps):
    for m in maps:
        if hasattr(m, "wcs"):
            return ndmap(arr, m.wcs)
    return arr


def geometry(pos, res, proj="car"):
    """enmap.geometry(pos=[[dec0,ra0],[dec1,ra1]], res, proj='car') as called at
    maps.py:1490.  CAR: crval=[mean(ra),0]; cdelt=+-res with the sign of
    pos[1]-pos[0] per axis; crpix such that pos[0] is the outer corner of pixel
    (0,0); shape = round(|pixel coordinate of pos[1]|) (numpy round-half-even).
    Known answer: +-5deg at 1' -> (600,600), cdelt=[1/60,1/60], crpix=[300.5,300.5]
    (tutorials/demo-grf.ipynb:97)."""
    assert proj == "car"
    pos = np.asarray(pos, dtype=np.float64) / degree  # -> degrees, rows [dec,ra]
    res = np.zeros(2) + np.asarray(res, dtype=np.float64) / degree
    posx = pos[:, ::-1]                                  # [x(ra), y(dec)] order
    mid = posx.mean(0)
    crval = np.array([mid[0], 0.0])
    cdelt = res.copy()
    cdelt[posx[1] < posx[0]] *= -1
    crpix = np.array([1.0, 1.0])
    off = (posx[0] - crval) / cdelt + crpix - 1 + 0.5    # 0-based pixel of pos[0], +0.5
    crpix = crpix - off
    faredge = (posx[1] - crval) / cdelt + crpix - 1      # 0-based pixel of pos[1]
    shape = tuple(int(v) for v in np.round(np.abs(faredge[::-1])))
    return shape, FlatWCS(cdelt, crval, crpix)


def _dec_edges(shape, wcs):
    ny = shape[-2]
    y = np.array([-0.5, ny - 0.5])
    dec = wcs.crval[1] + (y + 1 - wcs.crpix[1]) * wcs.cdelt[1]
    return dec * degree


def extent(shape, wcs, signed=False, method="cylindrical"):
    """enmap.extent.  'cylindrical' (modern pixell default for separable CAR):
    [Ny*cdelt_y, Nx*cdelt_x*(sin d2 - sin d1)/(d2-d1)] radians; 'intermediate':
    plain |cdelt|*shape."""
    ext = np.array([shape[-2], shape[-1]], dtype=np.float64) * wcs.cdelt[::-1] * degree
    if method == "cylindrical":
        d1, d2 = _dec_edges(shape, wcs)
        ext[1] *= (np.sin(d2) - np.sin(d1)) / (d2 - d1)
    elif method != "intermediate":
        raise ValueError(method)
    return ext if signed else np.abs(ext)


def area(shape, wcs, method="cylindrical"):
    """enmap.area as used at maps.py:1567,1605: sky area in steradians, equal to
    the product of the extents."""
    return float(np.prod(extent(shape, wcs, method=method)))


def pixsize(shape, wcs, method="cylindrical"):
    return area(shape, wcs, method) / np.prod(shape[-2:])


def laxes(shape, wcs, method="cylindrical"):
    """enmap.laxes: ly = 2pi fftfreq(Ny, dy), lx = 2pi fftfreq(Nx, dx), with the
    signed pixel steps dy,dx = extent/shape (maps.py:1939)."""
    step = extent(shape, wcs, signed=True, method=method) / np.array(shape[-2:])
    ly = np.fft.fftfreq(shape[-2], step[0]) * 2 * np.pi
    lx = np.fft.fftfreq(shape[-1], step[1]) * 2 * np.pi
    return ly, lx


def lmap(shape, wcs, method="cylindrical"):
    ly, lx = laxes(shape, wcs, method)
    out = np.empty((2,) + tuple(shape[-2:]))
    out[0] = ly[:, None]
    out[1] = lx[None, :]
    return ndmap(out, wcs)


def modlmap(shape, wcs, method="cylindrical"):
    """enmap.modlmap = sum(lmap**2,0)**0.5 (maps.py:1374,1938)."""
    return ndmap(np.sum(lmap(shape, wcs, method) ** 2, 0) ** 0.5, wcs)


def fft(emap, normalize=True, shape=None, wcs=None):
    """enmap.fft: forward c2c FFT over the last two axes.  normalize=True ->
    unitary (x Npix^-1/2); 'phys' additionally x pixsize^1/2; False -> raw
    (maps.py:1613 uses normalize=False; maps.py:1578 the default)."""
    res = scipy.fft.fft2(np.asarray(emap), axes=(-2, -1))
    if normalize:
        res = res / np.prod(res.shape[-2:]) ** 0.5
    if normalize in ("phy", "phys", "physical"):
        res = res * pixsize(res.shape, emap.wcs) ** 0.5
    return samewcs(res, emap)


def ifft(emap, normalize=True):
    """enmap.ifft: backward c2c; normalize=True -> unitary (maps.py:1585)."""
    a = np.asarray(emap)
    res = scipy.fft.ifft2(a, axes=(-2, -1)) * np.prod(a.shape[-2:])  # raw backward
    if normalize:
        res = res / np.prod(a.shape[-2:]) ** 0.5
    if normalize in ("phy", "phys", "physical"):
        res = res / pixsize(a.shape, emap.wcs) ** 0.5
    return samewcs(res, emap)


def raw_fft(a):
    """pixell.fft.fft(a, axes=[-2,-1]): raw forward c2c (maps.py:1923)."""
    return scipy.fft.fft2(np.asarray(a), axes=(-2, -1))


def raw_ifft(a, normalize=False):
    """pixell.fft.ifft(a, axes=[-2,-1], normalize): raw backward, /Npix if
    normalize (maps.py:1633, 1923)."""
    a = np.asarray(a)
    res = scipy.fft.ifft2(a, axes=(-2, -1))
    if not normalize:
        res = res * np.prod(a.shape[-2:])
    return res


def rand_gauss(shape, wcs):
    return ndmap(np.random.standard_normal(shape), wcs)


def rand_gauss_harm(shape, wcs):
    """enmap.rand_gauss_harm (maps.py:1578): real block drawn first, then the
    imaginary block, from numpy's global legacy RandomState."""
    return ndmap(np.random.standard_normal(shape) + 1j * np.random.standard_normal(shape), wcs)


def map_mul(mat, vec):
    """enmap.map_mul (maps.py:1579,1615): per-pixel matrix.vector over the leading
    component axes; <=3-D mat is plain broadcasting; 2-D vec is one component."""
    mat = np.asarray(mat)
    v = np.asarray(vec)
    if mat.ndim < 4:
        return samewcs(mat * v, vec)
    if v.ndim == 2:
        return samewcs(np.einsum("abyx,byx->ayx", mat, v[None])[0], vec)
    return samewcs(np.einsum("abyx,byx->ayx", mat, v), vec)


def eigpow(A, e):
    """pixell.utils.eigpow on axes [0,1] of a (n,n,...) real symmetric stack:
    eigen-decompose, raise eigenvalues to e; for non-integer e negative (and
    relatively tiny) eigenvalues are zeroed."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    if n == 1:
        a = A[0, 0]
        with np.errstate(invalid="ignore", divide="ignore"):
            out = np.where(a > 0, np.abs(a) ** e, 0.0) if e != int(e) or e < 0 else a ** e
        return out[None, None]
    M = np.moveaxis(A, (0, 1), (-2, -1))
    E, V = np.linalg.eigh(M)
    if e != int(e) or e < 0:
        emax = np.max(np.abs(E), -1, keepdims=True)
        bad = (E < emax * np.finfo(np.float64).resolution * 100) | (E < np.finfo(np.float64).tiny * 1e4)
        with np.errstate(invalid="ignore", divide="ignore"):
            Ee = np.where(bad, 0.0, np.abs(E) ** e)
    else:
        Ee = E ** e
    res = np.einsum("...ik,...k,...jk->...ij", V, Ee, V)
    return np.moveaxis(res, (-2, -1), (0, 1))


def multi_pow(mat, exp):
    """enmap.multi_pow(mat, exp, axes=[0,1]) (maps.py:1571)."""
    return samewcs(eigpow(mat, exp), mat)


def _convolute_sym(a, b):
    sa = np.concatenate([a, a[:, -2:0:-1]], -1)
    sb = np.concatenate([b, b[:, -2:0:-1]], -1)
    fa = scipy.fft.rfft(sa, axis=-1)
    fb = scipy.fft.rfft(sb, axis=-1)
    sa = scipy.fft.irfft(fa * fb, n=sa.shape[-1], axis=-1)
    return sa[:, : a.shape[-1]]


def smooth_spectrum(ps, kernel="gauss", weight="mode", width=1.0):
    """enmap.smooth_spectrum: sum(p W (*) K)/(W (*) K) along l with Gaussian K and
    W = l^2, convolution on the even-mirrored array."""
    ps = np.asarray(ps, dtype=np.float64)
    pflat = ps.reshape(-1, ps.shape[-1])
    nspec, nl = pflat.shape
    l = np.arange(nl, dtype=np.float64)
    K = np.zeros((nspec, nl))
    if kernel == "gauss":
        K[:] = np.exp(-0.5 * (l / width) ** 2)
    elif kernel == "step":
        K[:, : int(width)] = 1
    else:
        raise ValueError(kernel)
    W = np.zeros((nspec, nl))
    if weight == "mode":
        W[:] = l[None, :] ** 2
    elif weight == "uniform":
        W[:] = 1
    else:
        raise ValueError(weight)
    pWK = _convolute_sym(pflat * W, K)
    WK = _convolute_sym(W, K)
    with np.errstate(invalid="ignore", divide="ignore"):
        res = pWK / WK
    return res.reshape(ps.shape)


def spec2flat_1d(shape, wcs, cov, exp=1.0, smooth="auto", method="cylindrical"):
    """The 1-D half of spec2flat: smoothing, x Npix/area, matrix power, non-finite
    -> 0.  Returns (ncomp,ncomp,nl)."""
    cov = np.array(cov, dtype=np.float64)
    if cov.ndim == 1:
        cov = cov[None, None]
    ls = np.asarray(modlmap(shape, wcs, method))
    if smooth == "auto":
        smooth = 0.5 * (ls[1, 0] + ls[0, 1]) / 3.41
    if smooth and smooth > 0:
        cov = smooth_spectrum(cov, kernel="gauss", weight="mode", width=smooth)
    cov = cov * np.prod(shape[-2:]) / area(shape, wcs, method)
    if exp != 1.0:
        cov = eigpow(cov, exp)
    cov[~np.isfinite(cov)] = 0
    return cov


def spec2flat(shape, wcs, cov, exp=1.0, mode="constant", smooth="auto", method="cylindrical"):
    """enmap.spec2flat(shape,wcs,cov,exp,mode='constant',smooth='auto')
    (maps.py:1573): 1-D treatment above, then order-1 interpolation of each
    (i,j) spectrum at modlmap, zero outside the tabulated range.  Result is
    (ncomp,ncomp,Ny,Nx) even for 2-D shape."""
    assert mode == "constant"
    cov1 = spec2flat_1d(shape, wcs, cov, exp, smooth, method)
    ls = np.asarray(modlmap(shape, wcs, method))
    ncomp = cov1.shape[0]
    out = np.zeros((ncomp, ncomp) + ls.shape)
    for i in range(ncomp):
        for j in range(ncomp):
            out[i, j] = scipy.ndimage.map_coordinates(cov1[i, j], ls[None], order=1, mode="constant", cval=0.0)
    return ndmap(out, wcs)


def queb_rotmat(lm, inverse=False, iau=False, spin=2):
    """enmap.queb_rotmat (maps.py:1607): a = sgn*spin*atan2(-lx, ly), sgn=+1 for
    iau else -1; [[c,-s],[s,c]], s -> -s for the inverse; [E;B] = R [Q;U]."""
    lm = np.asarray(lm)
    sgn = 1 if iau else -1
    a = sgn * spin * np.arctan2(-lm[1], lm[0])
    c, s = np.cos(a), np.sin(a)
    if inverse:
        s = -s
    return np.array([[c, -s], [s, c]])


def harm2map(kmap, iau=False, method="cylindrical"):
    """enmap.harm2map (maps.py:1587): rotate components 1,2 (E,B)->(Q,U) with the
    inverse queb_rotmat, unitary ifft, .real."""
    k = np.array(kmap, dtype=np.complex128)
    if k.ndim > 2 and k.shape[-3] > 1:
        rot = queb_rotmat(lmap(k.shape, kmap.wcs, method), inverse=True, iau=iau)
        k[..., -2:, :, :] = np.einsum("abyx,byx->ayx", rot, k[..., -2:, :, :])
    return ndmap(ifft(ndmap(k, kmap.wcs), normalize=True).real, kmap.wcs)
