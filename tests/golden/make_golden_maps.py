"""Golden vectors for the orphics-owned logic of the flat-sky path: the reference's OWN class and function
bodies -- maps.rect_geometry (maps.py:1472-1498), MapGen (1553-1587), FourierCalc (1594-1677), get_taper /
cosine_window (1873-1920), filter_map (1922), gauss_beam (1925), mask_kspace (1936-1948) -- are cut out of
/root/reference with ast (the module cannot be imported: pixell/healpy/matplotlib are absent) and executed
UNMODIFIED in a namespace whose `enmap`, `utils`, `fft`, `ifft` are the numpy oracle's restatement of pixell
(oracle/enmap_np.py).  This pins every line orphics itself owns on the path (seeding, the
order of operations, normalisations, the power2d matrix layout, the rotation, the taper arithmetic...); what
stays unpinned is pixell's own behaviour behind those calls.

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_maps.py
"""
import ast
import os
import sys
import types

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
from oracle import enmap_np as oenmap, theory as otheory  # noqa: E402

path = "/root/reference/orphics/maps.py"
names = ["rect_geometry", "MapGen", "FourierCalc", "get_taper", "cosine_window", "filter_map", "gauss_beam", "mask_kspace"]
tree = ast.parse(open(path).read())
keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
assert len(keep) == len(names)

# pixell stand-ins: the oracle's restatement, with the extra keyword arguments the reference passes
enmap = types.SimpleNamespace(**{k: getattr(oenmap, k) for k in dir(oenmap) if not k.startswith("_")})
enmap.fft = lambda emap, nthread=0, normalize=True: oenmap.fft(emap, normalize=normalize)
enmap.laxes = lambda shape, wcs, oversample=1: oenmap.laxes(shape, wcs)
enmap.ones = lambda shape, wcs, dtype=float: oenmap.ndmap(np.ones(shape, dtype=dtype), wcs)
utils = types.SimpleNamespace(arcmin=oenmap.arcmin, degree=oenmap.degree)
ns = {"np": np, "enmap": enmap, "utils": utils,
      "fft": lambda a, axes=None: oenmap.raw_fft(a), "ifft": lambda a, axes=None, normalize=False: oenmap.raw_ifft(a, normalize=normalize)}
exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)

th = otheory.load_theory()
out = {}
for pol in (False, True):
    tag = "IQU" if pol else "T"
    shape, wcs = ns["rect_geometry"](width_arcmin=96 * 2.0, px_res_arcmin=2.0, height_arcmin=64 * 2.0, pol=pol)
    assert tuple(shape[-2:]) == (64, 96)
    modl = np.asarray(oenmap.modlmap(shape, wcs))
    ps = otheory.power_from_theory(np.arange(0, modl.max() + 1, 1.0), th, lensed=True, pol=pol)
    mg = ns["MapGen"](shape, wcs, ps)
    fc = ns["FourierCalc"](shape, wcs)
    m1, m2 = mg.get_map(seed=11), mg.get_map(seed=12)
    out[f"{tag}_ps"] = ps
    out[f"{tag}_covsqrt"] = np.asarray(mg.covsqrt)
    out[f"{tag}_map11"], out[f"{tag}_map12"] = np.asarray(m1), np.asarray(m2)
    out[f"{tag}_harm11"] = np.asarray(mg.get_map(seed=11, harm=True))
    if pol:
        out["IQU_map11_scalar"] = np.asarray(mg.get_map(seed=11, scalar=True))
        out["IQU_map11_iau"] = np.asarray(mg.get_map(seed=11, iau=True))
    p2d, k1, k2 = fc.power2d(m1)
    out[f"{tag}_p2d"], out[f"{tag}_k1"] = np.asarray(p2d), np.asarray(k1)
    px, kx1, kx2 = fc.power2d(m1, m2)
    out[f"{tag}_p2d_cross"] = np.asarray(px)
    out[f"{tag}_p2d_pix"] = np.asarray(fc.power2d(m1, pixel_units=True)[0])
    if pol:
        out["IQU_p2d_skip"] = np.asarray(fc.power2d(m1, skip_cross=True)[0])
        out["IQU_p2d_norot"] = np.asarray(fc.power2d(m1, rot=False)[0])
        out["IQU_teb_unitary"] = np.asarray(fc.iqu2teb(m1))
    else:
        f1, kk = fc.f1power(m2, k1)
        out["T_f1power"], out["T_ifft"], out["T_fft"] = np.asarray(f1), np.asarray(fc.ifft(k1)), np.asarray(fc.fft(m1))
        out["T_normfact"] = np.array(fc.normfact)
        taper, w2 = ns["get_taper"](shape, wcs)
        out["T_taper"], out["T_w2"] = np.asarray(taper), np.array(w2)
        out["T_taper_weight"] = np.asarray(ns["get_taper"](shape, wcs, taper_percent=20.0, pad_percent=5.0, weight=np.abs(np.asarray(m1)))[0])
        out["T_window_odd"] = ns["cosine_window"](37, 51, lenApodY=5, lenApodX=9, padY=2, padX=0)
        out["T_beam"] = ns["gauss_beam"](modl, 1.5)
        out["T_mask_l"] = np.asarray(ns["mask_kspace"](shape, wcs, lmin=300, lmax=2000))
        out["T_mask_xy"] = np.asarray(ns["mask_kspace"](shape, wcs, lxcut=90, lycut=50, lmax=4000))
        out["T_filtered"] = np.asarray(ns["filter_map"](m1, out["T_beam"] * out["T_mask_l"]))
np.savez_compressed(os.path.join(here, "maps_refbody.npz"), **out)
print({k: v.shape for k, v in out.items()})
