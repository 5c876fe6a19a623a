"""Golden vectors for the split / cross-spectrum callers: the reference's OWN function
bodies -- maps.split_calc (maps.py:2295-2332), maps.noise_from_splits (maps.py:2337-2411) and
lensing.SplitLensing (lensing.py:959-1003) -- are cut out of /root/reference with ast (the modules cannot
be imported: pixell/healpy/matplotlib are absent) and executed UNMODIFIED in a namespace whose
FourierCalc / enmap / qest are the numpy oracle's.  What is pinned is therefore the composition logic these
callers own (which pairs, which normalisations, the float32 cast, the never-taken rotation branch); the
FFT / f2power / estimator underneath are the oracle's.

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_callers.py
"""
import ast
import os
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
from oracle import enmap_np as oenmap, maps_np as omaps, qe_np, theory as otheory  # noqa: E402

REF = "/root/reference/orphics"


def cut(path, names):
    """compile the named top-level functions/classes of a reference source file, nothing else"""
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(keep) == len(names), (names, [n.name for n in keep])
    return compile(ast.Module(body=keep, type_ignores=[]), path, "exec")


class _MapsNS:          # what "maps." means inside the reference bodies
    FourierCalc = omaps.FourierCalc


ns_maps = {"np": np, "enmap": oenmap, "FourierCalc": omaps.FourierCalc, "maps": _MapsNS}
exec(cut(os.path.join(REF, "maps.py"), ["split_calc", "noise_from_splits"]), ns_maps)
ns_lens = {"np": np, "maps": _MapsNS}
exec(cut(os.path.join(REF, "lensing.py"), ["SplitLensing"]), ns_lens)


class _FC(omaps.FourierCalc):
    """the oracle's FourierCalc, tolerant of the nthread kwarg the reference passes (maps.py:2371)"""

    def iqu2teb(self, emap, nthread=0, normalize=True, rot=True):
        return super().iqu2teb(emap, normalize=normalize, rot=rot)


rng = np.random.RandomState(77)
out = {}
shape, wcs = omaps.rect_geometry(width_arcmin=48 * 2.0, px_res_arcmin=2.0, height_arcmin=32 * 2.0)
assert tuple(shape) == (32, 48)
fc = _FC(shape, wcs)
n = 4
sig = rng.standard_normal(shape)
isplits = oenmap.ndmap(np.stack([np.fft.fft2(sig + rng.standard_normal(shape)) for _ in range(n)]), wcs)
jsplits = oenmap.ndmap(np.stack([np.fft.fft2(0.5 * sig + rng.standard_normal(shape)) for _ in range(n)]), wcs)
ico, jco = isplits.mean(0), jsplits.mean(0)
out.update(sc_isplits=np.asarray(isplits), sc_jsplits=np.asarray(jsplits))
for alt in (True, False):
    t, c, nz = ns_maps["split_calc"](isplits, jsplits, ico, jco, fourier_calc=fc, alt=alt)
    out.update({f"sc_total_alt{int(alt)}": np.asarray(t), f"sc_cross_alt{int(alt)}": np.asarray(c), f"sc_noise_alt{int(alt)}": np.asarray(nz)})

for ncomp in (1, 3):
    sh = (ncomp,) + tuple(shape) if ncomp > 1 else tuple(shape)
    splits = oenmap.ndmap(rng.standard_normal((n,) + sh) + rng.standard_normal(sh), wcs)
    fcn = _FC(sh, wcs)
    for do_cross in (True, False):
        noise, cteb = ns_maps["noise_from_splits"](splits, fourier_calc=fcn, do_cross=do_cross)
        out[f"nfs_splits_c{ncomp}"] = np.asarray(splits)
        out[f"nfs_noise_c{ncomp}_x{int(do_cross)}"] = np.asarray(noise)
        if do_cross:
            out[f"nfs_crossteb_c{ncomp}"] = np.asarray(cteb)

# SplitLensing.cross_estimator with the oracle's qest on a 64 x 64, 4' patch
so, wo = omaps.rect_geometry(width_arcmin=64 * 4.0, px_res_arcmin=4.0)
th = otheory.load_theory()
modl = np.asarray(oenmap.modlmap(so, wo))
kw = dict(noise2d=np.zeros(so) + (10.0 * np.pi / 180 / 60) ** 2, beam2d=omaps.gauss_beam(modl, 5.0),
          kmask=np.asarray(omaps.mask_kspace(so, wo, lmin=200, lmax=2000)), kmask_K=np.asarray(omaps.mask_kspace(so, wo, lmin=50, lmax=2500)),
          unlensed_equals_lensed=True)
q = qe_np.qest(so, wo, th, **kw)
sl = ns_lens["SplitLensing"].__new__(ns_lens["SplitLensing"])
sl.fc, sl.qest, sl.est = omaps.FourierCalc(so, wo), q, "TT"
sigl = rng.standard_normal(so) * 60
ks = np.stack([np.fft.fft2(sigl + 20 * rng.standard_normal(so)) for _ in range(4)])
out.update(sl_ksplits=ks, sl_cross_estimator=np.asarray(sl.cross_estimator(ks)), sl_AL=np.asarray(q.N.AL["TT"]))
np.savez_compressed(os.path.join(here, "split_callers.npz"), **out)
print({k: v.shape for k, v in out.items()})
