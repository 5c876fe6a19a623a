"""ORACLE (test infrastructure) -- numpy restatement of the orphics.maps hot path.

The logic orphics owns is pinned by goldens made from the reference's own bodies (tests/golden/maps_refbody.npz,
split_callers.npz, ilc.npz); PARITY UNPINNED for what goes through pixell underneath (oracle/enmap_np.py, see
oracle/__init__.py).  Each function cites the /root/reference/orphics/maps.py lines it follows.
"""
import numpy as np

from . import enmap_np as enmap


def rect_geometry(width_arcmin=None, width_deg=None, px_res_arcmin=0.5, proj="car", pol=False,
                  height_deg=None, height_arcmin=None, xoffset_degree=0.0, yoffset_degree=0.0):
    """maps.py:1472-1498: box of the given width/height centred on the offsets."""
    if width_deg is not None:
        width_arcmin = 60.0 * width_deg
    if height_deg is not None:
        height_arcmin = 60.0 * height_deg
    hw = width_arcmin / 2.0
    vw = hw if height_arcmin is None else height_arcmin / 2.0
    am, dg = enmap.arcmin, enmap.degree
    pos = [[-vw * am + yoffset_degree * dg, -hw * am + xoffset_degree * dg],
           [vw * am + yoffset_degree * dg, hw * am + xoffset_degree * dg]]
    shape, wcs = enmap.geometry(pos=pos, res=px_res_arcmin * am, proj=proj)
    if pol:
        shape = (3,) + shape
    return shape, wcs


class MapGen:
    """maps.py:1553-1587."""

    def __init__(self, shape, wcs, cov=None, covsqrt=None, pixel_units=False, smooth="auto", method="cylindrical"):
        self.shape, self.wcs, self.method = shape, wcs, method
        if covsqrt is not None:                                  # maps.py:1563-1564
            self.covsqrt = np.asarray(covsqrt)
            return
        cov = np.asarray(cov)
        assert cov.ndim >= 3                                     # maps.py:1562
        if cov.ndim == 4:                                        # maps.py:1566-1571
            if not pixel_units:
                cov = cov * np.prod(shape[-2:]) / enmap.area(shape, wcs, method)
            self.covsqrt = enmap.multi_pow(cov, 0.5)
        else:                                                    # maps.py:1573
            self.covsqrt = enmap.spec2flat(shape, wcs, cov, 0.5, mode="constant", smooth=smooth, method=method)

    def kmap_from_noise(self, rand):
        """maps.py:1579: per-pixel covsqrt . rand."""
        return enmap.ndmap(enmap.map_mul(self.covsqrt, rand), self.wcs)

    def map_from_noise(self, rand, scalar=False, iau=False, harm=False):
        """maps.py:1579-1587 given the noise realisation ``rand`` (complex,
        self.shape)."""
        kmap = self.kmap_from_noise(rand)
        if harm:
            return kmap
        if scalar:
            return enmap.ndmap(enmap.ifft(kmap).real, self.wcs)  # maps.py:1585
        return enmap.harm2map(kmap, iau=iau, method=self.method)  # maps.py:1587

    def get_map(self, seed=None, scalar=False, iau=False, real=False, harm=False):
        """maps.py:1576-1587: numpy global legacy RNG, real block then imaginary."""
        if seed is not None:
            np.random.seed(seed)
        if real:
            rand = enmap.fft(enmap.rand_gauss(self.shape, self.wcs))
        else:
            rand = enmap.rand_gauss_harm(self.shape, self.wcs)
        return self.map_from_noise(rand, scalar=scalar, iau=iau, harm=harm)


class FourierCalc:
    """maps.py:1594-1677."""

    def __init__(self, shape, wcs, iau=False, method="cylindrical"):
        self.shape, self.wcs, self.method = shape, wcs, method
        self.normfact = enmap.area(shape, wcs, method) / np.prod(shape[-2:]) ** 2.0     # maps.py:1605
        if len(shape) > 2 and shape[-3] > 1:
            self.rot = enmap.queb_rotmat(enmap.lmap(shape, wcs, method), iau=iau)        # maps.py:1607

    def iqu2teb(self, emap, normalize=True, rot=True):
        """maps.py:1609-1617."""
        k = np.array(enmap.fft(emap, normalize=normalize))
        if k.ndim > 2 and k.shape[-3] > 1 and rot:
            k[..., -2:, :, :] = np.einsum("abyx,byx->ayx", self.rot, k[..., -2:, :, :])
        return enmap.ndmap(k, self.wcs)

    def f2power(self, kmap1, kmap2, pixel_units=False):
        """maps.py:1620-1624."""
        norm = 1.0 if pixel_units else self.normfact
        return np.real(np.conjugate(kmap1) * kmap2) * norm

    def f1power(self, map1, kmap2, pixel_units=False):
        """maps.py:1626-1630."""
        k1 = self.iqu2teb(map1, normalize=False)
        return self.f2power(k1, kmap2, pixel_units), k1

    def ifft(self, kmap):
        """maps.py:1632-1633: raw backward /Npix."""
        return enmap.ndmap(enmap.raw_ifft(kmap, normalize=True), self.wcs)

    def fft(self, emap):
        """maps.py:1635-1636."""
        return enmap.fft(emap, normalize=False)

    def power2d(self, emap=None, emap2=None, pixel_units=False, skip_cross=False, rot=True,
                kmap=None, kmap2=None, dtype=None):
        """maps.py:1639-1677."""
        if kmap is not None:
            l1 = kmap
            ndim = np.ndim(kmap)
        else:
            l1 = self.iqu2teb(emap, normalize=False, rot=rot)
            ndim = np.ndim(emap)
        ncomp = l1.shape[-3] if ndim > 2 else 1
        if kmap2 is not None:
            l2 = kmap2
        else:
            l2 = self.iqu2teb(emap2, normalize=False, rot=rot) if emap2 is not None else l1
        assert l1.shape == l2.shape
        if ndim > 2 and ncomp > 1:
            ret = np.zeros((ncomp, ncomp) + l1.shape[-2:], dtype=dtype)
            for i in range(ncomp):
                ret[i, i] = self.f2power(l1[i], l2[i], pixel_units)
            if not skip_cross:
                for i in range(ncomp):
                    for j in range(i + 1, ncomp):
                        ret[i, j] = self.f2power(l1[i], l2[j], pixel_units)
                        ret[j, i] = ret[i, j]
            return ret, l1, l2
        if l1.ndim > 2:
            l1 = l1[0]
        if l2.ndim > 2:
            l2 = l2[0]
        return (enmap.ndmap(self.f2power(l1, l2, pixel_units), self.wcs),
                enmap.ndmap(l1, self.wcs), enmap.ndmap(l2, self.wcs))


def cosine_window(Ny, Nx, lenApodY=30, lenApodX=30, padY=0, padX=0):
    """maps.py:1893-1920: separable raised-cosine edge taper with zero padding."""
    def axis(n, lap, pad):
        w = np.ones(n)
        i = np.arange(n)
        if lap > 0:
            r = i.astype(float) - pad
            s = i <= (lap + pad)
            w[s] = 0.5 * (1 - np.cos(-np.pi * r[s] / lap))
            s = i >= ((n - 1) - lap - pad)
            r = ((n - 1) - i - pad).astype(float)
            w[s] = 0.5 * (1 - np.cos(-np.pi * r[s] / lap))
        return w
    wx = axis(Nx, lenApodX, padX)
    wy = axis(Ny, lenApodY, padY)
    win = np.ones((Ny, Nx)) * wx[None, :]
    if lenApodY > 0:
        # the reference multiplies only the selected rows (maps.py:1911-1915); rows
        # outside the selection have wy == 1 so a plain product is identical.
        win = win * wy[:, None]
    win[0:padY, :] = 0
    win[:, 0:padX] = 0
    win[Ny - padY:, :] = 0
    win[:, Nx - padX:] = 0
    return win


def get_taper(shape, wcs, taper_percent=12.0, pad_percent=3.0, weight=None):
    """maps.py:1873-1879."""
    Ny, Nx = shape[-2:]
    if weight is None:
        weight = np.ones(shape[-2:])
    m = min(Ny, Nx)
    taper = cosine_window(Ny, Nx, lenApodY=int(taper_percent * m / 100.0), lenApodX=int(taper_percent * m / 100.0),
                          padY=int(pad_percent * m / 100.0), padX=int(pad_percent * m / 100.0)) * weight
    return enmap.ndmap(taper, wcs), np.mean(taper ** 2.0)


def filter_map(imap, kfilter):
    """maps.py:1922-1923."""
    return enmap.ndmap(np.real(enmap.raw_ifft(enmap.raw_fft(imap) * kfilter, normalize=True)), getattr(imap, "wcs", None))


def gauss_beam(ell, fwhm):
    """maps.py:1925-1927 (fwhm in arcmin)."""
    tht = np.deg2rad(fwhm / 60.0)
    return np.exp(-(tht ** 2.0) * (ell ** 2.0) / (16.0 * np.log(2.0)))


def mask_kspace(shape, wcs, lxcut=None, lycut=None, lmin=None, lmax=None, method="cylindrical"):
    """maps.py:1936-1948: integer mask, strict inequalities for the kept band."""
    out = np.ones(shape[-2:], dtype=int)
    if lmin is not None or lmax is not None:
        modl = np.asarray(enmap.modlmap(shape, wcs, method))
    if lxcut is not None or lycut is not None:
        ly, lx = enmap.laxes(shape, wcs, method)
    if lmin is not None:
        out[modl <= lmin] = 0
    if lmax is not None:
        out[modl >= lmax] = 0
    if lxcut is not None:
        out[:, np.abs(lx) < lxcut] = 0
    if lycut is not None:
        out[np.abs(ly) < lycut, :] = 0
    return enmap.ndmap(out, wcs)


def binned_power(imap, bin_edges, binner, fc, imap2=None, mask=1):
    """maps.py:1350-1361."""
    p2d, _, _ = fc.power2d(imap * mask, imap2 * mask if imap2 is not None else None)
    cents, p1d = binner.bin(p2d)
    return cents, p1d / np.mean(mask ** 2.0)


def split_calc(isplits, jsplits, icoadd, jcoadd, fourier_calc=None, alt=True):
    """maps.py:2295-2332: (total, crosses, noise) 2-D power from the Fourier transforms of splits
    and of their coadds.  alt=True: noise = sum_i <(a_i - A)*, (b_i - B)> / ((1 - 1/n) n^2),
    signal = total - noise; alt=False: signal = mean of the i != j cross spectra, noise = total - signal."""
    fc = fourier_calc
    assert np.ndim(isplits) == 3
    ni, nj = np.shape(isplits)[0], np.shape(jsplits)[0]
    total = fc.f2power(icoadd, jcoadd)
    if alt:
        assert ni == nj
        acc = 0.0
        for a, b in zip(isplits, jsplits):
            acc = acc + fc.f2power(a - icoadd, b - jcoadd)
        noise = acc / ((1.0 - 1.0 / ni) * ni ** 2)
        crosses = total - noise
    else:
        acc, cnt = 0.0, 0.0
        for i in range(ni):
            for j in range(nj):
                if i != j:
                    acc = acc + fc.f2power(isplits[i], jsplits[j])
                    cnt += 1.0
        crosses = acc / cnt
        noise = total - crosses
    return total, crosses, noise


def noise_from_splits(splits, fourier_calc=None, do_cross=True):
    """maps.py:2337-2411: noise model (auto - cross)/nsplits of the I,Q,U components from
    (nsplits, ncomp, Ny, Nx) splits (cast to float32 as the reference does, maps.py:2355), and the mean
    T,E,B cross spectrum of the i < j split pairs.  Spectra are power2d's (ncomp, ncomp, Ny, Nx) matrices
    (upper triangle computed, lower mirrored) for ncomp > 1, a 2-D map for ncomp = 1."""
    wcs = getattr(splits, "wcs", None)
    if wcs is None:
        wcs = splits[0].wcs
    splits = np.asarray(splits).astype(np.float32)
    assert splits.ndim in (3, 4)
    if splits.ndim == 3:
        splits = splits[:, None]
    n, ncomp = splits.shape[:2]
    fc = fourier_calc if fourier_calc is not None else FourierCalc(splits.shape[-3:] if do_cross else splits.shape[-2:], wcs)
    if do_cross:
        assert ncomp in (1, 3)
    ks = [fc.iqu2teb(enmap.ndmap(s, wcs), normalize=False, rot=False) for s in splits]
    kteb = None
    if do_cross:
        kteb = []
        for k in ks:
            r = np.array(k)
            # maps.py:2359 reads ndim AFTER a 3-D input has been promoted to 4-D, so the condition
            # "ndim==3 and ncomp==3" of maps.py:2379 is never true: the Q,U -> E,B rotation is never
            # applied and "cross_teb" is the I,Q,U cross spectrum.  Kept as written (bug-compatible).
            kteb.append(r)
    auto = 0.0
    for k in ks:
        auto = auto + fc.power2d(kmap=k)[0]
    auto = auto / n
    npairs = n * (n - 1) / 2
    cross, cross_teb = 0.0, (0.0 if do_cross else None)
    for i in range(n):
        for j in range(i + 1, n):
            cross = cross + fc.power2d(kmap=ks[i], kmap2=ks[j])[0]
            if do_cross:
                cross_teb = cross_teb + fc.power2d(kmap=kteb[i], kmap2=kteb[j])[0]
    cross = cross / npairs
    if do_cross:
        cross_teb = cross_teb / npairs
    return (auto - cross) / n, cross_teb


# ---- Fourier-space internal linear combination (maps.py:1952-2050): per-pixel small linear algebra
def _nan_to_num(x):
    return np.nan_to_num(x)


def ilc_comb_a_b(response_a, response_b, cinv):
    """a^T Cinv b per pixel (maps.py:2046-2049): sum_l a_l (sum_k b_k Cinv[k,l]), NaN -> 0, +-inf -> +-max."""
    inner = np.tensordot(np.asarray(response_b), np.asarray(cinv), axes=(0, 0))
    return _nan_to_num(np.tensordot(np.asarray(response_a), inner, axes=(0, 0)))


def ilc_map_term(kmaps, cinv, response):
    """response^T Cinv kmaps per pixel (maps.py:2042-2044)."""
    ck = np.sum(np.asarray(cinv) * np.asarray(kmaps)[None], axis=1)
    return np.tensordot(np.asarray(response), ck, axes=(0, 0))


def silc_noise(cinv, response=None):
    """maps.py:2020-2023."""
    response = np.ones(np.shape(cinv)[0]) if response is None else response
    with np.errstate(divide="ignore", invalid="ignore"):
        return _nan_to_num(1.0 / ilc_comb_a_b(response, response, cinv))


def silc(kmaps, cinv, response=None):
    """Standard ILC estimate (maps.py:1952-1974)."""
    response = np.ones(np.shape(cinv)[0]) if response is None else response
    return ilc_map_term(kmaps, cinv, response) * silc_noise(cinv, response)


def cilc(kmaps, cinv, response_a, response_b):
    """Constrained ILC: component a with component b projected out (maps.py:1976-2005)."""
    brb = ilc_comb_a_b(response_b, response_b, cinv)
    arb = ilc_comb_a_b(response_a, response_b, cinv)
    ara = ilc_comb_a_b(response_a, response_a, cinv)
    arM = ilc_map_term(kmaps, cinv, response_a)
    brM = ilc_map_term(kmaps, cinv, response_b)
    with np.errstate(divide="ignore", invalid="ignore"):
        return _nan_to_num((brb * arM - arb * brM) / (ara * brb - arb ** 2.0))


def cilc_noise(cinv, response_a, response_b):
    """maps.py:2025-2039."""
    brb = ilc_comb_a_b(response_b, response_b, cinv)
    ara = ilc_comb_a_b(response_a, response_a, cinv)
    arb = ilc_comb_a_b(response_a, response_b, cinv)
    bra = ilc_comb_a_b(response_b, response_a, cinv)
    with np.errstate(divide="ignore", invalid="ignore"):
        return _nan_to_num((brb ** 2.0 * ara + arb ** 2.0 * brb - brb * arb * arb - arb * brb * bra) / (ara * brb - arb ** 2.0) ** 2.0)
