"""orphics_b200 -- the flat-sky Fourier hot path of msyriac/orphics on NVIDIA B200.

Drop-in for that path only: ``maps.MapGen``, ``maps.FourierCalc``, ``stats.bin2D``,
``lensing.qest`` (and the few helpers around them) with the reference's signatures,
calling through ctypes into ``_lib/liborphx.so`` (hand-written CUDA for sm_100a + cuFFT).
There is no CPU fallback: importing works without a GPU, computing does not.
"""
from . import _capi  # noqa: F401  (loads liborphx.so; raises if it has not been built)
from . import enmap, maps, stats, mpi, cosmology, lensing  # noqa: F401

__version__ = "0.1.0"
