"""ctypes binding of liborphx.so (include/orphx.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
usable, every compute entry point raises.  Importing this module only loads the
library (possible without a GPU); the first compute call needs the device.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ORPHX_LIB: alternative build of the same library (kernel-tuning experiments, tools/sweep_variants.sh)
LIB_PATH = os.environ.get("ORPHX_LIB") or os.path.join(_HERE, "_lib", "liborphx.so")

OX_F64, OX_F32 = 0, 1
OX_HOST, OX_DEVICE = 0, 1
NOISE_HOST, NOISE_PHILOX, NOISE_PHILOX_HERMITIAN = 0, 1, 2
FLAG_ROT, FLAG_SKIP_CROSS, FLAG_PIXEL_UNITS, FLAG_IAU, FLAG_MASK_NAN, FLAG_HARM, FLAG_UNITARY = 1, 2, 4, 8, 16, 32, 64
FLAG_KEEP_MAPS = 128
QE_TT, QE_EB = 0, 1

NOISE_MODES = {"host": NOISE_HOST, "numpy": NOISE_HOST, "philox": NOISE_PHILOX,
               "philox_hermitian": NOISE_PHILOX_HERMITIAN, "hermitian": NOISE_PHILOX_HERMITIAN}


class OrphxError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise OrphxError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C orphics_b200/csrc`. orphics_b200 has no CPU fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()

_vp, _i, _ll, _d, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_size_t
_pd = C.POINTER(C.c_double)
_pll = C.POINTER(C.c_longlong)
_pvp = C.POINTER(C.c_void_p)

# name -> argtypes; every function returns int status except ox_last_error / ox_abi_version
SIGNATURES = {
    "ox_abi_version": [],
    "ox_device_count": [C.POINTER(_i)],
    "ox_set_device": [_i],
    "ox_get_device": [C.POINTER(_i)],
    "ox_device_name": [C.c_char_p, _sz],
    "ox_synchronize": [],
    "ox_set_stream": [_vp],
    "ox_malloc": [_pvp, _sz],
    "ox_free": [_vp],
    "ox_malloc_pooled": [_pvp, _sz],
    "ox_free_pooled": [_vp],
    "ox_map_op": [_i, _vp, _vp, _d, _ll, _ll, _i, _vp],
    "ox_memset": [_vp, _i, _sz],
    "ox_host_alloc": [_pvp, _sz],
    "ox_host_free": [_vp],
    "ox_memcpy_h2d": [_vp, _vp, _sz],
    "ox_memcpy_d2h": [_vp, _vp, _sz],
    "ox_memcpy_d2d": [_vp, _vp, _sz],
    "ox_mem_info": [C.POINTER(_sz), C.POINTER(_sz)],
    "ox_timer_create": [_pvp],
    "ox_timer_start": [_vp],
    "ox_timer_stop": [_vp],
    "ox_timer_elapsed_ms": [_vp, C.POINTER(C.c_float)],
    "ox_timer_destroy": [_vp],
    "ox_launch_count": [_pll],
    "ox_flush_l2": [],
    "ox_profile_begin": [],
    "ox_profile_end": [C.POINTER(_i)],
    "ox_profile_stage": [_i, C.c_char_p, _sz, C.POINTER(C.c_float)],
    "ox_geometry_create": [_i, _i, _vp, _vp, _d, _pvp],
    "ox_geometry_destroy": [_vp],
    "ox_geometry_modlmap": [_vp, _vp, _i],
    "ox_geometry_rotmat": [_vp, _i, _vp, _i],
    "ox_geometry_mask_kspace": [_vp, _d, _d, _d, _d, _vp, _i],
    "ox_geometry_interp_spec": [_vp, _vp, _i, _i, _vp, _i],
    "ox_binner_create": [_vp, _i, _ll, _vp, _i, _pvp],
    "ox_binner_create_geom": [_vp, _vp, _i, _pvp],
    "ox_binner_destroy": [_vp],
    "ox_binner_digitized": [_vp, _vp],
    "ox_binner_counts": [_vp, _vp],
    "ox_binner_bin": [_vp, _vp, _i, _i, _ll, _vp, _i, _vp, _vp, _i],
    "ox_simplan_create": [_vp, _i, _vp, _i, _i, _i, _pvp],
    "ox_simplan_destroy": [_vp],
    "ox_sim_generate": [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _i],
    "ox_powerplan_create": [_vp, _i, _i, _i, _pvp],
    "ox_powerplan_destroy": [_vp],
    "ox_power_fft": [_vp, _vp, _i, _i, _i, _vp, _i],
    "ox_power_ifft": [_vp, _vp, _i, _i, _vp, _i],
    "ox_power_f2power": [_vp, _vp, _vp, _i, _ll, _i, _vp, _i],
    "ox_power2d": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i],
    "ox_power_bin": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _i],
    "ox_pipeline_create": [_vp, _vp, _vp, _vp, _i, _pvp],
    "ox_pipeline_destroy": [_vp],
    "ox_pipeline_run": [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _i],
    "ox_pipeline_path": [_vp, C.POINTER(_i)],
    "ox_pipeline_maps": [_vp, _pvp],
    "ox_pipeline_profile": [_vp, _vp, _i, _i, _i, C.POINTER(C.c_float)],
    "ox_pipeline_stats": [_vp, _pvp, C.POINTER(_i)],
    "ox_pipeline_stats_reset": [_vp],
    "ox_qeplan_create": [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _pvp],
    "ox_qeplan_destroy": [_vp],
    "ox_qe_reconstruct": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i],
    "ox_qe_meanfield": [_vp, _pvp, _pll],
    "ox_comm_unique_id": [_vp, _sz],
    "ox_comm_create": [_i, _i, _vp, _sz, _pvp],
    "ox_comm_destroy": [_vp],
    "ox_comm_info": [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)],
    "ox_comm_allreduce_f64": [_vp, _vp, _ll],
    "ox_pipeline_allreduce": [_vp, _vp],
    "ox_qe_meanfield_allreduce": [_vp, _vp],
    "ox_qe_meanfield_reset": [_vp],
    "ox_qe_path": [_vp],
    "ox_fft_c2c": [_vp, _vp, _i, _i, _i, _d, _vp, _i],
    "ox_power_filter": [_vp, _vp, _i, _i, _vp, _i, _vp, _i],
    "ox_power_filter_complex": [_vp, _vp, _i, _i, _vp, _i, _vp, _i],
    "ox_split_calc": [_vp, _vp, _vp, _vp, _i, _i, _i, C.c_longlong, _d, _i, _vp, _vp, _vp, _i],
    "ox_noise_from_splits": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i],
    "ox_split_lensing_combine": [_vp, _i, _i, _i, C.c_longlong, _d, _vp, _i],
    "ox_ilc": [_vp, _vp, _vp, _vp, _i, C.c_longlong, _i, _i, _vp, _i],
    "ox_multi_pow": [_vp, _i, C.c_longlong, _d, _i, _vp, _i],
    "ox_plane_symmetry": [_vp, _vp, _vp],
    "ox_qe_filter": [_vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _vp],
    "ox_qe_norm": [_vp, _i, _vp, _vp, _vp, _vp, _d, _d, _vp, _vp],
    "ox_lensplan_create": [_vp, _d, _d, _i, _pvp],
    "ox_lensplan_destroy": [_vp],
    "ox_lens_kappa_to_phi": [_vp, _vp, _i, _vp, _i],
    "ox_lens_set_phi": [_vp, _vp, _i],
    "ox_lens_alpha": [_vp, _vp, _i],
    "ox_lens_taylens": [_vp, _vp, _i, _i, _i, _vp, _i],
    "ox_lens_displace": [_vp, _vp, _i, _i, _vp, _i],
}

for _name, _args in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = _i
lib.ox_last_error.argtypes = []
lib.ox_last_error.restype = C.c_char_p


def check(status):
    if status != 0:
        raise OrphxError(f"liborphx status {status}: {lib.ox_last_error().decode(errors='replace')}")


_device_ready = False


def require_device():
    """Raise unless a CUDA device is usable (no CPU fallback exists)."""
    global _device_ready
    if _device_ready:
        return
    n = C.c_int(0)
    st = lib.ox_device_count(C.byref(n))
    if st != 0 or n.value < 1:
        raise OrphxError("orphics_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback: "
                         + lib.ox_last_error().decode(errors="replace"))
    _device_ready = True


def ptr(a):
    """void* of a numpy array (must be C-contiguous) or an int device address or None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return C.c_void_p(a.ctypes.data)


def np_dtype(dtype):
    return np.float32 if dtype == OX_F32 else np.float64


def np_cdtype(dtype):
    return np.complex64 if dtype == OX_F32 else np.complex128


def ox_dtype(dtype):
    """Map a numpy dtype / None / 'f32' to OX_F64 / OX_F32."""
    if dtype is None:
        return OX_F64
    if dtype in (OX_F64, OX_F32) and not isinstance(dtype, type):
        return dtype
    dt = np.dtype(dtype)
    if dt in (np.dtype(np.float32), np.dtype(np.complex64)):
        return OX_F32
    if dt in (np.dtype(np.float64), np.dtype(np.complex128)):
        return OX_F64
    raise ValueError(f"unsupported dtype {dtype}")


class DeviceBuffer:
    """Owning handle of a cudaMalloc'd buffer (used by the batched APIs and bench)."""

    def __init__(self, nbytes):
        require_device()
        p = C.c_void_p()
        check(lib.ox_malloc(C.byref(p), nbytes))
        self.ptr = p.value
        self.nbytes = nbytes

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(lib.ox_memcpy_h2d(C.c_void_p(self.ptr), ptr(arr), arr.nbytes))
        return self

    def download(self, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(lib.ox_memcpy_d2h(ptr(out), C.c_void_p(self.ptr), out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib.ox_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """numpy view of cudaHostAlloc'd memory."""

    def __init__(self, shape, dtype):
        require_device()
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib.ox_host_alloc(C.byref(p), max(nbytes, 1)))
        self._p = p.value
        buf = (C.c_char * max(nbytes, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p:
            self.array = None
            lib.ox_host_free(C.c_void_p(self._p))
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Timer:
    def __init__(self):
        require_device()
        self._t = C.c_void_p()
        check(lib.ox_timer_create(C.byref(self._t)))

    def start(self):
        check(lib.ox_timer_start(self._t))

    def stop(self):
        check(lib.ox_timer_stop(self._t))

    def elapsed_ms(self):
        ms = C.c_float(0)
        check(lib.ox_timer_elapsed_ms(self._t, C.byref(ms)))
        return float(ms.value)

    def __del__(self):
        try:
            lib.ox_timer_destroy(self._t)
        except Exception:
            pass


class StageProfile:
    """with StageProfile() as p: ...calls...; p.stages -> [(name, ms)] from CUDA events between the library's
    stage marks (ox_profile_*)."""

    def __enter__(self):
        check(lib.ox_profile_begin())
        self.stages = []
        return self

    def __exit__(self, *exc):
        n = C.c_int(0)
        check(lib.ox_profile_end(C.byref(n)))
        buf = C.create_string_buffer(64)
        for i in range(n.value):
            ms = C.c_float(0)
            check(lib.ox_profile_stage(i, buf, 64, C.byref(ms)))
            self.stages.append((buf.value.decode(), float(ms.value)))
        return False

    def totals(self):
        out = {}
        for name, ms in self.stages:
            out[name] = out.get(name, 0.0) + ms
        return out


def launch_count():
    n = C.c_longlong(0)
    check(lib.ox_launch_count(C.byref(n)))
    return int(n.value)


def synchronize():
    check(lib.ox_synchronize())


def set_device(dev):
    require_device()
    check(lib.ox_set_device(int(dev)))


def device_name():
    require_device()
    buf = C.create_string_buffer(256)
    check(lib.ox_device_name(buf, 256))
    return buf.value.decode()
