"""Debug helper: the estimator with the 2-row TMA c2r pass (ORPHX_KB_R2=1) against the one-tile kernel on the same inputs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orphics_b200 import maps, lensing, cosmology
ny, nx = int(sys.argv[1]), int(sys.argv[2])
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 2
shape, wcs = maps.rect_geometry(width_arcmin=nx * 0.5, px_res_arcmin=0.5, height_arcmin=ny * 0.5)
assert tuple(shape) == (ny, nx), shape
th = cosmology.default_theory()
modl = maps.Geometry.get(shape, wcs).modlmap()
kw = dict(noise2d=np.zeros(shape) + (1.0 * np.pi / 180 / 60) ** 2, beam2d=maps.gauss_beam(modl, 1.5),
          kmask=maps.mask_kspace(shape, wcs, lmin=300, lmax=2000), kmask_K=maps.mask_kspace(shape, wcs, lmin=20, lmax=3500),
          unlensed_equals_lensed=True, max_batch=nb)
q = lensing.qest(shape, wcs, th, **kw)
T = np.random.RandomState(1).standard_normal((nb,) + tuple(shape)) * 50
out = {}
for rep in range(3):
    for r2 in ("0", "1"):
        os.environ["ORPHX_KB_R2"] = r2
        out[r2] = np.asarray(q.kappa_from_maps("TT", T, returnFt=True))
    d = np.abs(out["1"] - out["0"]).max() / np.abs(out["0"]).max()
    print(ny, nx, nb, q.path("TT"), "R2 vs one-tile kernel rel diff", d, flush=True)
