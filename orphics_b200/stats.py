"""orphics.stats hot-path mirror: bin2D (stats.py:782-811) on the device, and the
Statistics reduce triple (stats.py:918-1232) over torch.distributed.

Same names, arguments and quirks as the reference:
  * bin2D(modrmap, bin_edges) with attributes centers/cents/digitized/bin_edges/modrmap;
  * .bin(data2d, weights=None, err=False, get_count=False, mask_nan=False);
  * np.bincount(...)[1:-1] trimming (stats.py:796-797), including its behaviour when
    no pixel lies above the last edge.
The device returns raw per-slot sums/counts (ox_binner_bin); trimming and the final
division of these tiny arrays happen here.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import lib, check, ptr


class bin2D(object):
    def __init__(self, modrmap, bin_edges, geometry=None):
        """geometry: optional enmap.Geometry -- when given, the slot index is derived
        on the device from the geometry's Fourier axes (modrmap must be its modlmap)
        and the fused FourierCalc.power2d + bin path becomes available."""
        _capi.require_device()
        bin_edges = np.asarray(bin_edges)
        self.centers = (bin_edges[1:] + bin_edges[:-1]) / 2.
        self.cents = self.centers  # backwards compatibility (stats.py:785)
        self.bin_edges = bin_edges
        self.modrmap = modrmap
        self._edges64 = np.ascontiguousarray(bin_edges, dtype=np.float64)
        h = C.c_void_p()
        if geometry is not None:
            self._n = geometry.npix
            # the slot index is derived on the device from the geometry's own |l|: refuse any other modrmap
            # (a real-space radius map, a masked or edited modlmap) instead of silently binning by |l|
            if modrmap is not None:
                m = np.asarray(modrmap)
                if m.shape != tuple(geometry.shape) or not np.array_equal(m, geometry.modlmap()):
                    raise ValueError("bin2D(geometry=...): modrmap must be the geometry's modlmap(); "
                                     "drop geometry= to bin by an arbitrary modrmap")
            check(lib.ox_binner_create_geom(geometry.handle, ptr(self._edges64), self._edges64.size, C.byref(h)))
        else:
            m = np.ascontiguousarray(np.asarray(modrmap), dtype=np.float64)
            self._n = m.size
            check(lib.ox_binner_create(ptr(m), _capi.OX_HOST, m.size, ptr(self._edges64), self._edges64.size, C.byref(h)))
        self.handle = h
        self.geometry = geometry
        self._digitized = None
        self.nslots = self._edges64.size + 1
        cnt = np.empty(self.nslots, dtype=np.int64)
        check(lib.ox_binner_counts(self.handle, ptr(cnt)))
        self.slot_counts = cnt  # == np.bincount(digitized, minlength=len(edges)+1)
        #: number of bandpowers the reference returns: np.bincount(digitized)[1:-1] is only max(digitized)+1 long
        #: (stats.py:796-797), so when no pixel lies above the last edge the last occupied bin is trimmed too.
        #: The fixed-shape fused outputs (binned_power_batch, SimPipeline) are cut to this length on the host.
        self.trimmed_nbins = int(self._trim(cnt, cnt).size)

    @property
    def digitized(self):
        """int64 slot index per pixel, identical to np.digitize(modrmap.ravel(), edges, right=True)."""
        if self._digitized is None:
            out = np.empty(self._n, dtype=np.int64)
            check(lib.ox_binner_digitized(self.handle, ptr(out)))
            self._digitized = out
        return self._digitized

    def _raw(self, data, weights, flags):
        from .enmap import devmap
        if not isinstance(data, devmap):
            data = np.asarray(data)
        # float32 data stay float32 on the device unless weights are given: the reference's weights are float64
        # (np.bincount promotes), so weighted sums are formed in float64
        dt = _capi.OX_F32 if (data.dtype == np.float32 and weights is None) else _capi.OX_F64
        if data.size % self._n != 0:
            raise ValueError(f"data size {data.size} is not a multiple of the binner's {self._n} pixels")
        nmaps = data.size // self._n
        sums = np.empty((nmaps, self.nslots), dtype=np.float64)
        cnts = np.empty((nmaps, self.nslots), dtype=np.float64)
        if isinstance(data, devmap) and weights is None and data.dtype == _capi.np_dtype(dt):
            # a device-resident map (FourierCalc.power2d's result) is binned where it is
            check(lib.ox_binner_bin(self.handle, C.c_void_p(data.ptr), dt, _capi.OX_DEVICE, nmaps, None, flags, ptr(sums), ptr(cnts),
                                    _capi.OX_HOST))
            return sums, cnts
        d = np.ascontiguousarray(np.asarray(data), dtype=_capi.np_dtype(dt))
        w = None
        if weights is not None:
            w = np.ascontiguousarray(np.broadcast_to(np.asarray(weights), np.shape(self.modrmap)), dtype=d.dtype)
        check(lib.ox_binner_bin(self.handle, ptr(d), dt, _capi.OX_HOST, nmaps, ptr(w), flags, ptr(sums), ptr(cnts), _capi.OX_HOST))
        return sums, cnts

    @staticmethod
    def _trim(occupied_src, arr):
        """np.bincount(x)[1:-1]: the bincount is only max(x)+1 long."""
        occ = np.nonzero(occupied_src)[0]
        length = int(occ[-1]) + 1 if occ.size else 0
        return arr[:length][1:-1]

    def bin(self, data2d, weights=None, err=False, get_count=False, mask_nan=False):
        from .enmap import devmap
        if not isinstance(data2d, devmap) or err or weights is not None:
            data2d = np.asarray(data2d)
        if data2d.size != self._n:
            raise ValueError("data2d does not match the binner's modrmap size; use bin_batch for stacks")
        if weights is None:
            sums, cnts = self._raw(data2d, None, _capi.FLAG_MASK_NAN if mask_nan else 0)
            icount = np.rint(cnts[0]).astype(np.int64)
            count = self._trim(icount, icount)
            with np.errstate(invalid="ignore", divide="ignore"):
                res = self._trim(icount, sums[0]) / count
            if err:
                # stats.py:798-801, including its slot/bin off-by-one; evaluated on the host
                # from the device-computed means (not on the north-star path)
                dig = self.digitized
                flat = data2d.reshape(-1)
                meanmap = np.zeros(flat.size)
                for i in range(self.centers.size):
                    meanmap[dig == i] = res[i]
                dev2 = (flat - meanmap) ** 2.
                s2, _ = self._raw(dev2.reshape(data2d.shape), None, _capi.FLAG_MASK_NAN if mask_nan else 0)
                with np.errstate(invalid="ignore", divide="ignore"):
                    std = np.sqrt(self._trim(icount, s2[0]) / (count - 1) / count)
        else:
            sums, cnts = self._raw(data2d, np.asarray(weights), 0)
            count = self._trim(self.slot_counts, cnts[0])
            with np.errstate(invalid="ignore", divide="ignore"):
                res = self._trim(self.slot_counts, sums[0]) / count
        if get_count:
            assert not (err)  # as the reference
            return self.centers, res, count
        if err:
            assert not (get_count)
            return self.centers, res, std
        return self.centers, res

    def bin_batch(self, data, mask_nan=False):
        """Bin a stack (..., Ny, Nx) in one device pass; returns (centers, res[..., nbins])."""
        from .enmap import devmap
        if not isinstance(data, devmap):
            data = np.asarray(data)
        lead = data.shape[:-np.ndim(self.modrmap)] if np.ndim(self.modrmap) else data.shape[:-1]
        sums, cnts = self._raw(data, None, _capi.FLAG_MASK_NAN if mask_nan else 0)
        out = []
        for s, c in zip(sums, cnts):
            ic = np.rint(c).astype(np.int64)
            with np.errstate(invalid="ignore", divide="ignore"):
                out.append(self._trim(ic, s) / self._trim(ic, ic))
        return self.centers, np.array(out).reshape(lead + (-1,))

    def __del__(self):
        try:
            lib.ox_binner_destroy(self.handle)
        except Exception:
            pass


def bin_in_annuli(data2d, modrmap, bin_edges):
    """stats.py:853-855."""
    binner = bin2D(modrmap, bin_edges)
    return binner.bin(data2d)


class Statistics:
    """The reduce semantics of orphics.stats.Statistics (stats.py:918-1419): per label
    N, SUM x, SUM x x^T (stats mode) or K, SUM arr (stack mode); ``allreduce`` sums them
    over the ranks of a torch.distributed process group (NCCL on GPUs, gloo on CPU) instead
    of mpi4py -- as ONE collective of a packed buffer, whatever the number of labels.  ``comm`` is a torch.distributed process group, ``True`` for the default
    group, or None for single-process use."""

    def __init__(self, comm=None, dtype=np.float64, nccl=None):
        """nccl: optional mpi.NcclComm -- the sums then travel through ox_comm_allreduce_f64 (NCCL issued by the C
        library); the torch.distributed group only exchanges the label metadata."""
        self.comm = comm
        self.nccl = nccl
        self.dtype = np.dtype(dtype)
        self._n, self._sum, self._cross = {}, {}, {}
        self._k, self._stack = {}, {}
        self._reduced = False

    @property
    def mpi_enabled(self):
        return self.comm is not None

    def _stats_label(self, label, d):
        if label in self._stack:
            raise ValueError(f"Label {label!r} already used in stack mode.")
        if label not in self._sum:
            self._n[label] = 0
            self._sum[label] = np.zeros(d, dtype=self.dtype)
            self._cross[label] = np.zeros((d, d), dtype=self.dtype)
        elif self._sum[label].shape[0] != d:
            raise ValueError(f"Stats dim mismatch for {label!r}: {self._sum[label].shape[0]} vs {d}")

    def _stack_label(self, label, shape):
        if label in self._sum:
            raise ValueError(f"Label {label!r} already used in stats mode.")
        if label not in self._stack:
            self._k[label] = 0
            self._stack[label] = np.zeros(shape, dtype=self.dtype)
        elif self._stack[label].shape != tuple(shape):
            raise ValueError(f"Stack shape mismatch for {label!r}: {self._stack[label].shape} vs {tuple(shape)}")

    def add(self, label, x):
        x = np.asarray(x, dtype=self.dtype).ravel()
        self._stats_label(label, x.shape[0])
        self._n[label] += 1
        self._sum[label] += x
        self._cross[label] += np.outer(x, x)

    def extend(self, label, X):
        X = np.asarray(list(X) if not hasattr(X, "shape") else X, dtype=self.dtype)
        if X.ndim == 1:
            return self.add(label, X)
        if X.ndim != 2:
            raise ValueError("X must be (m, d) or (d,).")
        self._stats_label(label, X.shape[1])
        self._n[label] += X.shape[0]
        self._sum[label] += X.sum(axis=0)
        self._cross[label] += X.T @ X

    def add_triple(self, label, n, s, c):
        """Merge a device-accumulated (N, SUM, CROSS) triple (ox_pipeline_stats)."""
        s = np.asarray(s, dtype=self.dtype)
        self._stats_label(label, s.shape[0])
        self._n[label] += int(n)
        self._sum[label] += s
        self._cross[label] += np.asarray(c, dtype=self.dtype)

    def add_stack(self, label, arr):
        A = np.asarray(arr, dtype=self.dtype)
        self._stack_label(label, A.shape)
        self._k[label] += 1
        self._stack[label] += A

    def allreduce(self):
        if self.mpi_enabled:
            import torch
            import torch.distributed as dist
            group = None if self.comm is True else self.comm
            ws = dist.get_world_size(group)
            local = {"stats": sorted((repr(l), l, int(v.shape[0])) for l, v in self._sum.items()),
                     "stack": sorted((repr(l), l, tuple(v.shape)) for l, v in self._stack.items())}
            gathered = [None] * ws
            dist.all_gather_object(gathered, local, group=group)
            stats_union, stack_union = {}, {}
            for entry in gathered:
                for _, lab, d in entry["stats"]:
                    if lab in stats_union and stats_union[lab] != d:
                        raise ValueError(f"Stats dim mismatch for {lab!r} across ranks.")
                    if lab in stack_union:
                        raise ValueError(f"Label {lab!r} used in stats and stack across ranks.")
                    stats_union[lab] = d
                for _, lab, shp in entry["stack"]:
                    if lab in stack_union and stack_union[lab] != tuple(shp):
                        raise ValueError(f"Stack shape mismatch for {lab!r} across ranks.")
                    if lab in stats_union:
                        raise ValueError(f"Label {lab!r} used in stats and stack across ranks.")
                    stack_union[lab] = tuple(shp)
            for lab, d in stats_union.items():
                self._stats_label(lab, d)
            for lab, shp in stack_union.items():
                self._stack_label(lab, shp)
            # ONE collective for everything (the reference issues 3 per stats label and 2 per stack label,
            # stats.py:1215-1217, 1227-1228): [N | SUM | CROSS] of every stats label and [K | SUM] of every stack label
            # packed into one float64 buffer (counts stay exact below 2^53)
            slabs, klabs = sorted(stats_union, key=repr), sorted(stack_union, key=repr)
            parts = []
            for lab in slabs:
                parts += [np.array([self._n[lab]], dtype=np.float64), np.asarray(self._sum[lab], dtype=np.float64).ravel(),
                          np.asarray(self._cross[lab], dtype=np.float64).ravel()]
            for lab in klabs:
                parts += [np.array([self._k[lab]], dtype=np.float64), np.asarray(self._stack[lab], dtype=np.float64).ravel()]
            packed = np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0)
            if packed.size:
                if self.nccl is not None:
                    # data plane in the C ABI: upload, ncclAllReduce issued by liborphx.so, download
                    buf = _capi.DeviceBuffer(packed.nbytes).upload(packed)
                    self.nccl.allreduce_f64(buf.ptr, packed.size)
                    packed = buf.download(packed.shape, np.float64)
                    buf.free()
                else:
                    backend = dist.get_backend(group)
                    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
                    t = torch.from_numpy(packed).to(dev)
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                    packed = t.cpu().numpy()
            o = 0
            for lab in slabs:
                d = stats_union[lab]
                self._n[lab] = int(round(packed[o]))
                self._sum[lab] = packed[o + 1:o + 1 + d].astype(self.dtype)
                self._cross[lab] = packed[o + 1 + d:o + 1 + d + d * d].reshape(d, d).astype(self.dtype)
                o += 1 + d + d * d
            for lab in klabs:
                shp = stack_union[lab]
                m = int(np.prod(shp, dtype=np.int64))
                self._k[lab] = int(round(packed[o]))
                self._stack[lab] = packed[o + 1:o + 1 + m].reshape(shp).astype(self.dtype)
                o += 1 + m
        self._reduced = True

    def _check(self):
        if not self._reduced:
            raise RuntimeError("Call .allreduce() before requesting global stats/stack.")

    def labels_stats(self):
        return list(self._sum.keys())

    def labels_stack(self):
        return list(self._stack.keys())

    def count(self, label):
        self._check()
        if label not in self._sum:
            raise KeyError(f"{label!r} is not a stats-mode label.")
        return self._n[label]

    def stack_count(self, label):
        self._check()
        if label not in self._stack:
            raise KeyError(f"{label!r} is not a stack-mode label.")
        return self._k[label]

    def mean(self, label):
        n = self.count(label)
        S = self._sum[label]
        return S / n if n > 0 else np.full(S.shape, np.nan, dtype=self.dtype)

    def cov(self, label, ddof=1):
        n = self.count(label)
        S, Cm = self._sum[label], self._cross[label]
        d = S.shape[0]
        if n <= ddof:
            return np.full((d, d), np.nan, dtype=self.dtype)
        return (Cm - np.outer(S, S) / n) / (n - ddof)

    def var(self, label, ddof=1):
        n = self.count(label)
        S, Cm = self._sum[label], self._cross[label]
        if n <= ddof:
            return np.full(S.shape[0], np.nan, dtype=self.dtype)
        return (np.diag(Cm) - S * S / n) / (n - ddof)

    def stack_sum(self, label):
        self._check()
        if label not in self._stack:
            raise KeyError(f"{label!r} is not a stack-mode label.")
        return self._stack[label]

    # checkpoint format of the reference (stats.py:1455-1530): .npz with keys
    # stats/<label>/{N,SUM,CROSS} and stack/<label>/SUM, written by the root rank only
    def save_reduced(self, path, compressed=False, root_rank=0):
        self._check()
        if self.mpi_enabled:
            import torch.distributed as dist
            group = None if self.comm is True else self.comm
            if dist.get_rank(group) != root_rank:
                return
        arrays = {}
        for lab in self._sum:
            arrays[f"stats/{lab}/N"] = np.array(self._n[lab], dtype=np.int64)
            arrays[f"stats/{lab}/SUM"] = self._sum[lab]
            arrays[f"stats/{lab}/CROSS"] = self._cross[lab]
        for lab in self._stack:
            arrays[f"stack/{lab}/SUM"] = self._stack[lab]
        (np.savez_compressed if compressed else np.savez)(path, **arrays)

    @classmethod
    def load_reduced(cls, path, comm=None, dtype=np.float64):
        data = np.load(path, allow_pickle=False)
        acc = cls(comm=comm, dtype=dtype)
        for key in data.files:
            kind, lab, what = key.split("/")
            if kind == "stats":
                if what == "N":
                    acc._n[lab] = int(data[key])
                elif what == "SUM":
                    acc._sum[lab] = np.array(data[key])
                elif what == "CROSS":
                    acc._cross[lab] = np.array(data[key])
            elif kind == "stack" and what == "SUM":
                acc._stack[lab] = np.array(data[key])
                acc._k.setdefault(lab, 0)     # the reference does not store the stack count either
        acc._reduced = True
        return acc
