"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/orphx.h declares; compute calls fail loudly without a device."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_gpu


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "orphx.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ox_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from orphics_b200 import _capi
    syms = header_symbols()
    assert len(syms) > 40
    for s in syms:
        assert hasattr(_capi.lib, s), f"{s} declared in include/orphx.h but not exported by liborphx.so"
    assert _capi.lib.ox_abi_version() == 2


def test_ctypes_signatures_cover_the_header():
    from orphics_b200 import _capi
    declared = set(header_symbols()) - {"ox_last_error"}
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)


def test_library_is_sm100a_native():
    from orphics_b200 import _capi
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


@pytest.mark.skipif(has_gpu(), reason="a GPU is present")
def test_compute_fails_loudly_without_device():
    import numpy as np
    from orphics_b200 import _capi, maps, stats
    with pytest.raises(_capi.OrphxError):
        _capi.require_device()
    with pytest.raises(_capi.OrphxError):
        stats.bin2D(np.ones((8, 8)), np.array([0.0, 1.0, 2.0]))
    shape, wcs = maps.rect_geometry(width_arcmin=64.0, px_res_arcmin=2.0)
    with pytest.raises(_capi.OrphxError):
        maps.FourierCalc(shape, wcs)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "orphics_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                # the only scipy use allowed in the product is scipy.fft for the 1-D spectrum smoothing of the
                # set-up (maps._sym_convolve; numpy.fft before): no scipy compute on any per-pixel path
                uses = re.findall(r"scipy[.\w]*", src)
                assert all(u in ("scipy.fft",) for u in uses) and (not uses or fn == "maps.py"), (fn, uses)
