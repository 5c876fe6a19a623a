"""Set-up time of lensing.qest (filters + A_L, TT and EB) with the device set-up and with the numpy set-up of round 1
(ORPHX_QE_SETUP=host), at the BASELINE sizes.  Usage: python tools/bench_qe_setup.py [npix ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(npix, mode):
    os.environ["ORPHX_QE_SETUP"] = mode
    from orphics_b200 import maps, lensing, cosmology, _capi
    shape, wcs = maps.rect_geometry(width_arcmin=npix * 0.5, px_res_arcmin=0.5)
    th = cosmology.default_theory()
    g = maps.Geometry.get(shape, wcs)
    modl = g.modlmap()
    beam = maps.gauss_beam(modl, 1.5)
    n2d = np.zeros(shape) + (1.0 * np.pi / 180 / 60) ** 2
    tm = np.asarray(maps.mask_kspace(shape, wcs, lmin=300, lmax=2000))
    km = np.asarray(maps.mask_kspace(shape, wcs, lmin=20, lmax=3500))
    _capi.synchronize()
    t0 = time.perf_counter()
    q = lensing.qest(shape, wcs, th, noise2d=n2d, beam2d=beam, kmask=tm, noise2d_P=2 * n2d, kmask_P=tm, kmask_K=km, pol=True,
                     unlensed_equals_lensed=True)
    _capi.synchronize()
    dt = time.perf_counter() - t0
    chk = float(np.asarray(q.N.AL["TT"]).sum())
    del q
    return dt, chk


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096]
    out = {}
    for n in sizes:
        row = {}
        for mode in ("device", "host"):
            run(n, mode) if mode == "device" and n == sizes[0] else None      # warm the context / plans once
            dt, chk = run(n, mode)
            row[mode] = {"seconds": dt, "sum_AL_TT": chk}
        out[str(n)] = row
        print(n, row, flush=True)
    print(json.dumps({"qest_setup_seconds_TT_and_EB": out}))
