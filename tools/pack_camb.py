"""Pack the three CAMB text tables the hot path needs into one .npz of RAW columns
(no arithmetic applied), so that the product, the oracle and bench.py can load the
same theory spectra on the GPU box where /root/reference does not exist.

Run here:  python tools/pack_camb.py
Source:    /root/reference/data/cosmo2017_10K_acc3_{lensedCls,scalCls,lenspotentialCls}.dat
"""
import hashlib
import sys

import numpy as np

root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/cosmo2017_10K_acc3"
out = sys.argv[2] if len(sys.argv) > 2 else "orphics_b200/data/cosmo2017_10K_acc3.npz"
cols = {}
md5 = {}
for suffix, use in (("_lensedCls.dat", [0, 1, 2, 3, 4]), ("_scalCls.dat", [0, 1, 2, 3]),
                    ("_lenspotentialCls.dat", [0, 5])):
    fn = root + suffix
    cols[suffix[1:-4]] = np.loadtxt(fn, usecols=use)
    md5[suffix[1:-4]] = hashlib.md5(open(fn, "rb").read()).hexdigest()
np.savez_compressed(out, md5=np.array([f"{k}:{v}" for k, v in md5.items()]), **cols)
print(out, {k: v.shape for k, v in cols.items()}, md5)
