"""Golden vectors for the FFT-based lensing callers: the reference's OWN function bodies -- lensing.flat_taylens
(lensing.py:395-440), kappa_to_phi / kappa_to_fphi / fkappa_to_fphi (lensing.py:651-665) -- cut out of
/root/reference with ast and executed UNMODIFIED over the oracle's pixell stand-in (oracle/enmap_np.py).

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_lensing.py
"""
import ast
import os
import sys
import types

import numpy as np
from scipy.special import factorial

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
from oracle import enmap_np as oenmap, maps_np as omaps  # noqa: E402

path = "/root/reference/orphics/lensing.py"
names = ["flat_taylens", "kappa_to_phi", "kappa_to_fphi", "fkappa_to_fphi"]
tree = ast.parse(open(path).read())
keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
assert len(keep) == len(names)


class _Map(oenmap.ndmap):
    """the oracle's ndmap with the two pixell methods the bodies call on their arguments"""

    def lmap(self):
        return oenmap.lmap(self.shape, self.wcs)

    def modlmap(self):
        return oenmap.modlmap(self.shape, self.wcs)


enmap = types.SimpleNamespace(**{k: getattr(oenmap, k) for k in dir(oenmap) if not k.startswith("_")})
enmap.laxes = lambda shape, wcs, oversample=1: oenmap.laxes(shape, wcs)
enmap.pixshape = lambda shape, wcs: oenmap.extent(shape, wcs) / np.array(shape[-2:])
ns = {"np": np, "enmap": enmap, "factorial": factorial}
exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)

rng = np.random.RandomState(23)
shape, wcs = omaps.rect_geometry(width_arcmin=64 * 2.0, px_res_arcmin=2.0, height_arcmin=48 * 2.0)
assert tuple(shape) == (48, 64)
modl = np.asarray(oenmap.modlmap(shape, wcs))
kappa = _Map(rng.standard_normal(shape) * 0.05, wcs)
phi, fphi = ns["kappa_to_phi"](kappa, modl, return_fphi=True)
# a smooth potential with displacements of a few pixels, so that the integer shifts of Taylens are exercised
lp = np.exp(-0.5 * (modl / 300.0) ** 2)
phis = oenmap.ndmap(np.real(np.fft.ifft2(np.fft.fft2(rng.standard_normal(shape)) * lp)) * 3e-4, wcs)
imap = _Map(rng.standard_normal(shape) * 50, wcs)
out = {"kappa": np.asarray(kappa), "phi": np.asarray(phi), "fphi": np.asarray(fphi), "phis": np.asarray(phis), "imap": np.asarray(imap)}
for order in (2, 5):
    out[f"lensed_o{order}"] = np.asarray(ns["flat_taylens"](_Map(phis, wcs), imap, order))
out["fk2fp"] = ns["fkappa_to_fphi"](np.fft.fft2(np.asarray(kappa)), modl)
alphaX = np.real(oenmap.ifft(oenmap.ndmap(1j * np.asarray(oenmap.lmap(shape, wcs))[1] * np.asarray(oenmap.fft(phis, normalize="phys")), wcs), normalize="phys"))
out["max_shift_pixels"] = np.array(np.max(np.abs(alphaX)) / (oenmap.extent(shape, wcs)[1] / shape[1]))
np.savez_compressed(os.path.join(here, "lensing_refbody.npz"), **out)
print({k: v.shape for k, v in out.items()}, float(out["max_shift_pixels"]))
