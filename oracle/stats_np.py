"""ORACLE (test infrastructure) -- numpy restatement of orphics.stats.bin2D,
the Statistics reduce triple and orphics.mpi.mpi_distribute.

PINNED: these three are importable from the reference itself, and
tests/golden/make_golden.py ran the reference's own code
(PYTHONPATH=/root/reference) to produce tests/golden/bin2d_*.npz,
statistics_*.npz and mpi_distribute.npz, against which tests/test_oracle.py
checks this restatement bit for bit.
"""
import numpy as np


class bin2D:
    """stats.py:782-811."""

    def __init__(self, modrmap, bin_edges):
        bin_edges = np.asarray(bin_edges)
        self.centers = (bin_edges[1:] + bin_edges[:-1]) / 2.0          # stats.py:784
        self.cents = self.centers
        # stats.py:786: right=True => slot b <=> edges[b-1] < x <= edges[b]
        self.digitized = np.digitize(np.asarray(modrmap).reshape(-1), bin_edges, right=True)
        self.bin_edges = bin_edges
        self.modrmap = modrmap

    def bin(self, data2d, weights=None, err=False, get_count=False, mask_nan=False):
        data2d = np.asarray(data2d)
        flat = data2d.reshape(-1)
        if weights is None:
            keep = ~np.isnan(flat) if mask_nan else np.ones(flat.size, dtype=bool)   # stats.py:792-795
            dig = self.digitized[keep]
            count = np.bincount(dig)[1:-1]                                             # stats.py:796
            with np.errstate(invalid="ignore", divide="ignore"):
                res = np.bincount(dig, flat[keep])[1:-1] / count                       # stats.py:797
            if err:                                                                    # stats.py:798-801
                meanmap = np.zeros(flat.size)
                for i in range(self.centers.size):
                    meanmap[self.digitized == i] = res[i]      # reference's off-by-one kept as is
                d2 = ((data2d - meanmap.reshape(data2d.shape)) ** 2.0).reshape(-1)[keep]
                with np.errstate(invalid="ignore", divide="ignore"):
                    std = np.sqrt(np.bincount(dig, d2)[1:-1] / (count - 1) / count)
        else:
            w = np.asarray(weights).reshape(-1)
            count = np.bincount(self.digitized, w)[1:-1]                               # stats.py:803
            with np.errstate(invalid="ignore", divide="ignore"):
                res = np.bincount(self.digitized, (data2d * weights).reshape(-1))[1:-1] / count   # stats.py:804
        if get_count:
            assert not err
            return self.centers, res, count
        if err:
            return self.centers, res, std
        return self.centers, res


def bin_in_annuli(data2d, modrmap, bin_edges):
    """stats.py:853-855."""
    return bin2D(modrmap, bin_edges).bin(data2d)


def mpi_distribute(num_tasks, avail_cores, allow_empty=False):
    """mpi.py:78-91: contiguous split, the remainder goes to the LAST ranks."""
    if not allow_empty:
        assert avail_cores <= num_tasks
    min_each, rem = divmod(num_tasks, avail_cores)
    num_each = np.full(avail_cores, min_each)
    if rem > 0:
        num_each[-rem:] += 1
    ends = np.cumsum(num_each)
    starts = ends - num_each
    return num_each, [list(range(s, e)) for s, e in zip(starts, ends)]


class StatsTriple:
    """The reduce semantics of stats.Statistics (stats.py:1068-1232, 1311-1419):
    per label N, SUM x, SUM x x^T (stats mode) or K, SUM arr (stack mode);
    ``merge`` is what Allreduce(SUM) does across ranks."""

    def __init__(self):
        self.N, self.SUM, self.CROSS, self.K, self.STACK = {}, {}, {}, {}, {}

    def add(self, label, x):
        x = np.asarray(x, dtype=np.float64).ravel()
        if label not in self.N:
            self.N[label], self.SUM[label], self.CROSS[label] = 0, np.zeros(x.size), np.zeros((x.size, x.size))
        self.N[label] += 1
        self.SUM[label] += x
        self.CROSS[label] += np.outer(x, x)

    def add_stack(self, label, arr):
        a = np.asarray(arr, dtype=np.float64)
        if label not in self.K:
            self.K[label], self.STACK[label] = 0, np.zeros(a.shape)
        self.K[label] += 1
        self.STACK[label] += a

    @staticmethod
    def merge(parts):
        out = StatsTriple()
        for p in parts:
            for l in p.N:
                if l not in out.N:
                    out.N[l], out.SUM[l], out.CROSS[l] = 0, np.zeros_like(p.SUM[l]), np.zeros_like(p.CROSS[l])
                out.N[l] += p.N[l]
                out.SUM[l] = out.SUM[l] + p.SUM[l]
                out.CROSS[l] = out.CROSS[l] + p.CROSS[l]
            for l in p.K:
                if l not in out.K:
                    out.K[l], out.STACK[l] = 0, np.zeros_like(p.STACK[l])
                out.K[l] += p.K[l]
                out.STACK[l] = out.STACK[l] + p.STACK[l]
        return out

    def mean(self, label):
        return self.SUM[label] / self.N[label]                         # stats.py:1311-1333

    def cov(self, label, ddof=1):
        n, S, C = self.N[label], self.SUM[label], self.CROSS[label]   # stats.py:1335-1364
        return (C - np.outer(S, S) / n) / (n - ddof)
