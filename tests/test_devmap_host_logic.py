"""CPU test of the HOST logic of enmap.devmap (dispatch rules, views, item assignment, fallbacks to numpy): the device is
replaced by a stand-in for the six C entry points the class calls, with "device memory" kept in host buffers and
ox_map_op done by numpy.  The real kernels are covered by tests/test_gpu_devmap.py."""
import ctypes as C

import numpy as np
import pytest


class _FakeLib:
    """ox_malloc_pooled / ox_free_pooled / ox_memcpy_* / ox_map_op over host buffers."""

    def __init__(self):
        self.blocks = {}
        self.map_ops = 0

    def ox_malloc_pooled(self, pref, n):
        buf = (C.c_char * max(int(n), 1))()
        addr = C.addressof(buf)
        self.blocks[addr] = buf
        C.cast(pref, C.POINTER(C.c_void_p))[0] = addr
        return 0

    def ox_free_pooled(self, p):
        self.blocks.pop(p.value if isinstance(p, C.c_void_p) else int(p), None)
        return 0

    @staticmethod
    def _addr(p):
        return p.value if isinstance(p, C.c_void_p) else int(p)

    def ox_memcpy_h2d(self, dst, src, n):
        C.memmove(self._addr(dst), self._addr(src), int(n))
        return 0

    ox_memcpy_d2h = ox_memcpy_h2d
    ox_memcpy_d2d = ox_memcpy_h2d

    def ox_map_op(self, op, a, b, scalar, n, nb, kind, out):
        self.map_ops += 1
        adt = [np.float64, np.float32, np.complex128, np.complex64][kind]
        bdt = [np.float64, np.float32, np.float64, np.float32][kind]
        n, nb = int(n), int(nb)
        x = np.frombuffer((C.c_char * (n * np.dtype(adt).itemsize)).from_address(self._addr(a)), dtype=adt)
        if b is None:
            y = bdt(scalar.value if isinstance(scalar, C.c_double) else scalar)
        else:
            y = np.tile(np.frombuffer((C.c_char * (nb * np.dtype(bdt).itemsize)).from_address(self._addr(b)), dtype=bdt), n // nb)
        res = [lambda: x * y, lambda: x + y, lambda: x - y, lambda: x / y, lambda: y - x, lambda: y / x][op]()
        o = np.frombuffer((C.c_char * (n * np.dtype(adt).itemsize)).from_address(self._addr(out)), dtype=adt)
        o[:] = res.astype(adt)
        return 0


@pytest.fixture()
def dm(monkeypatch):
    from orphics_b200 import enmap, _capi
    fake = _FakeLib()
    monkeypatch.setattr(enmap, "lib", fake)
    monkeypatch.setattr(_capi, "require_device", lambda: None)
    return enmap, fake


def test_arithmetic_stays_on_the_device_where_numpy_rules_allow(dm):
    enmap, fake = dm
    rng = np.random.RandomState(0)
    h, w = rng.standard_normal((6, 8)), rng.uniform(0.5, 1.5, (6, 8))
    m, wd = enmap.devmap.from_host(h, "wcs"), enmap.devmap.from_host(w)
    cases = [(m * w, h * w), (w * m, w * h), (m * wd, h * w), (m + wd, h + w), (m - wd, h - w), (wd - m, w - h), (m / wd, h / w),
             (w / m, w / h), (m * 2.5, h * 2.5), (2.5 * m, 2.5 * h), (1.0 - m, 1.0 - h), (m + 1, h + 1), (np.multiply(w, m), w * h),
             (m * w[0], h * w[0])]
    for got, want in cases:
        assert isinstance(got, enmap.devmap) and got.wcs == "wcs" or isinstance(got, enmap.devmap)
        assert np.array_equal(np.asarray(got), want)
    assert fake.map_ops == len(cases)
    n0 = fake.map_ops
    # numpy's own rules (promotion, integer operands, general broadcasting, other ufuncs): host results, no device op
    for got, want in [(m * np.arange(8), h * np.arange(8)), (m ** 2, h ** 2), (-m, -h), (m > 0, h > 0), (np.sqrt(abs(m)), np.sqrt(abs(h))),
                      (m * w.astype(np.float32), h * w.astype(np.float32)), (m * (1 + 2j), h * (1 + 2j)), (m * w[:, :1], h * w[:, :1])]:
        assert not isinstance(got, enmap.devmap) and np.array_equal(got, want)
    assert fake.map_ops == n0
    # complex map (op) real filter
    k = rng.standard_normal((6, 8)) + 1j * rng.standard_normal((6, 8))
    kd = enmap.devmap.from_host(k)
    assert np.array_equal(np.asarray(kd * wd), k * w) and np.array_equal(np.asarray(kd / 2.0), k / 2.0)
    assert not isinstance(kd * kd, enmap.devmap) and np.array_equal(kd * kd, k * k)


def test_views_item_assignment_and_read_only_host_copy(dm):
    enmap, fake = dm
    h = np.arange(3 * 4 * 5, dtype=np.float64).reshape(3, 4, 5)
    m = enmap.devmap.from_host(h)
    m._host = None                                             # as a result of a device call: only the device copy exists
    v = m[1]
    assert isinstance(v, enmap.devmap) and v.shape == (4, 5) and np.array_equal(np.asarray(v), h[1]) and np.array_equal(np.asarray(m[-1]), h[2])
    assert np.array_equal(m[0, 1:3], h[0, 1:3]) and m.shape == h.shape and m.ndim == 3 and len(m) == 3 and m.size == 60
    assert [np.asarray(x).sum() for x in m] == [h[i].sum() for i in range(3)]
    assert np.array_equal(np.asarray(m.reshape(12, 5)), h.reshape(12, 5)) and m.mean() == h.mean() and np.array_equal(m.T, h.T)
    with pytest.raises(ValueError):
        np.asarray(m)[0, 0, 0] = 9.0                           # the cached host copy is read-only
    c = np.array(m)
    c[0, 0, 0] = 9.0
    assert np.asarray(m)[0, 0, 0] == 0.0
    v[0, :] = -1.0                                             # edits the view's host copy; never writes through into the parent
    want = h[1].copy()
    want[0, :] = -1.0
    assert np.array_equal(np.asarray(v * 1.0), want) and np.array_equal(np.asarray(m), h)
    m2 = m.copy()
    np.multiply(h, 2.0, out=m2)                                # ufunc out= into a devmap: host write, device copy follows
    assert np.array_equal(np.asarray(m2 + 0.0), 2.0 * h)
    with pytest.raises(IndexError):
        m[3]
    with pytest.raises(TypeError):
        hash(m)


def test_dense_theory_tables_reproduce_the_piecewise_linear_theory():
    """Host half of the estimator's device set-up (lensing._dense_theory_tables): the dense integer-l tables interpolate
    linearly to exactly what TheorySpectra.lCl / uCl return between the tabulated knots and above the last one; the first
    knot is reported so that the caller can refuse geometries with modes below it; foreign theory objects are refused."""
    from orphics_b200 import cosmology, lensing
    th = cosmology.default_theory()
    tabs, lo = lensing._dense_theory_tables(th, 21600.0)
    assert lo == 2.0 and set(tabs) == {w + k for w in "lu" for k in ("TT", "EE", "BB", "TE")}
    L = np.concatenate([np.linspace(2.0, 9500.0, 20011), [0.0, 8249.5, 8250.0, 8250.5, 12000.0]])
    for which, f in (("l", th.lCl), ("u", th.uCl)):
        for k in ("TT", "EE", "TE", "BB"):
            t = tabs[which + k]
            got = np.interp(L, np.arange(t.size, dtype=np.float64), t, left=0.0, right=0.0)
            want = f(k, L)
            assert np.max(np.abs(got - want)) <= 1e-15 * np.max(np.abs(want)) + 0.0, (which, k)

    class Other:
        def lCl(self, k, L):
            return L * 0

        uCl = lCl
    assert lensing._dense_theory_tables(Other(), 3000.0) is None
