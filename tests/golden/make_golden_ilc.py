"""Golden vectors for the Fourier-space ILC: maps.silc / cilc / silc_noise / cilc_noise
(maps.py:1952-2050) are pure numpy, so the reference's own functions are cut out of /root/reference with
ast and executed UNMODIFIED (the module itself cannot be imported: pixell/healpy are absent).

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ilc.py
"""
import ast
import os

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
names = ["silc", "cilc", "ilc_def_response", "ilc_index", "silc_noise", "cilc_noise", "ilc_map_term", "ilc_comb_a_b"]
path = "/root/reference/orphics/maps.py"
tree = ast.parse(open(path).read())
keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
assert len(keep) == len(names)
ns = {"np": np}
exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)

rng = np.random.RandomState(5)
out = {}
for tag, nfreq, trail in (("2d", 5, (24, 40)), ("1d", 3, (57,)), ("pair", 2, (16, 16))):
    npix = int(np.prod(trail))
    A = rng.standard_normal((npix, nfreq, nfreq))
    cov = A @ A.transpose(0, 2, 1) + 0.5 * np.eye(nfreq)              # SPD per pixel
    cinv = np.linalg.inv(cov).transpose(1, 2, 0).reshape((nfreq, nfreq) + trail).copy()
    kmaps = (rng.standard_normal((nfreq,) + trail) + 1j * rng.standard_normal((nfreq,) + trail)) * 30
    # a few degenerate pixels: zero matrix (1/0 -> inf -> max float), NaN entries, a rank-1 block
    flat = cinv.reshape(nfreq, nfreq, npix)
    flat[:, :, 0] = 0.0
    flat[:, :, 1] = np.nan
    flat[:, :, 2] = 1.0
    ra = rng.uniform(0.5, 1.5, nfreq)
    rb = rng.uniform(-1.0, 2.0, nfreq)
    with np.errstate(all="ignore"):
        out.update({f"{tag}_cinv": cinv, f"{tag}_kmaps": kmaps, f"{tag}_ra": ra, f"{tag}_rb": rb,
                    f"{tag}_silc": ns["silc"](kmaps, cinv), f"{tag}_silc_resp": ns["silc"](kmaps, cinv, ra),
                    f"{tag}_cilc": ns["cilc"](kmaps, cinv, ra, rb),
                    f"{tag}_silc_noise": ns["silc_noise"](cinv), f"{tag}_silc_noise_resp": ns["silc_noise"](cinv, ra),
                    f"{tag}_cilc_noise": ns["cilc_noise"](cinv, ra, rb)})
np.savez_compressed(os.path.join(here, "ilc.npz"), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()})
