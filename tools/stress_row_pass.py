"""Stress check of the persistent TMA row pass: many batches of the 2048^2 pipeline (221 tiles per CTA and launch) with the TMA
kernel and with the one-tile kernel on the same seeds; bandpowers must agree to rounding for every map (a single misread tile
moves a bandpower by ~1e-3).  Usage: python tools/stress_row_pass.py [nbatch] [batch]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orphics_b200 import maps, stats, cosmology  # noqa: E402

nbatch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
shape, wcs = maps.rect_geometry(width_arcmin=2048 * 0.5, px_res_arcmin=0.5)
g = maps.Geometry.get(shape, wcs)
ps = cosmology.power_from_theory(np.arange(0, g.modlmap().max() + 1, 1.0), cosmology.default_theory(), lensed=True, pol=False)
taper = np.asarray(maps.get_taper(shape, wcs)[0])
edges = np.arange(100, 3000, 40.0)
out = {}
for kb in ("legacy", "tma"):
    os.environ["ORPHX_KB"] = kb
    mg = maps.MapGen(shape, wcs, ps, noise="philox_hermitian", max_batch=B)
    fc = maps.FourierCalc(shape, wcs, max_batch=B)
    b = stats.bin2D(g.modlmap(), edges, geometry=g)
    pipe = maps.SimPipeline(mg, fc, b, window=taper)
    out[kb] = pipe.run(range(5000, 5000 + nbatch * B), keep_maps=True)
a, c = out["tma"], out["legacy"]
rel = np.abs(a - c) / np.abs(c)
worst = float(np.nanmax(rel))
bad = int((np.nanmax(rel.reshape(rel.shape[0], -1), axis=1) > 1e-12).sum())
print(f"{nbatch * B} maps: max relative bandpower difference TMA vs one-tile kernel {worst:.3e}; maps off by more than 1e-12: {bad}")
sys.exit(0 if bad == 0 else 1)
