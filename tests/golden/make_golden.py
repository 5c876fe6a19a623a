"""Generate golden vectors from the REFERENCE's own importable code.

Run in the build container only (the reference is not on the GPU box):
    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_golden.py

Uses orphics.stats.bin2D (stats.py:782-811), orphics.stats.Statistics
(stats.py:918-1419) with the single-rank fake communicator, and
orphics.mpi.mpi_distribute (mpi.py:78-91), all UNMODIFIED.  Inputs are seeded
and stored with the outputs so the fixtures are self-contained.
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
os.environ["DISABLE_MPI"] = "true"
from orphics import stats as rstats  # noqa: E402
from orphics import mpi as rmpi  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))


def lgrid(ny, nx, dy, dx):
    ly = np.fft.fftfreq(ny, dy) * 2 * np.pi
    lx = np.fft.fftfreq(nx, dx) * 2 * np.pi
    return np.sqrt(ly[:, None] ** 2 + lx[None, :] ** 2)


def bin2d_case(name, modr, edges, rng, with_nan=False):
    data = rng.standard_normal(modr.shape) * (1 + modr / modr.max())
    w = rng.uniform(0.5, 1.5, modr.shape)
    b = rstats.bin2D(modr, edges)
    out = dict(modrmap=modr, edges=edges, data=data, weights=w, digitized=b.digitized, centers=b.centers)
    with np.errstate(all="ignore"):
        c, r, n = b.bin(data, get_count=True)
        out.update(res=r, count=n)
        c, rw, nw = b.bin(data, weights=w, get_count=True)
        out.update(res_w=rw, count_w=nw)
        if with_nan:
            d2 = data.copy()
            d2[rng.uniform(size=d2.shape) < 0.05] = np.nan
            c, rn, nn = b.bin(d2, mask_nan=True, get_count=True)
            out.update(data_nan=d2, res_nan=rn, count_nan=nn)
        try:
            c, re_, se = b.bin(data, err=True)
            out.update(res_err=re_, std_err=se)
        except Exception as e:  # the reference's err branch breaks in the trim quirk case
            out.update(err_exception=np.array(repr(e)))
    np.savez_compressed(os.path.join(here, f"bin2d_{name}.npz"), **out)
    print(name, modr.shape, len(edges), r.shape, n.sum())


rng = np.random.RandomState(1234)
# 1. annular binning of a Fourier modlmap, pixels beyond the last edge exist
m = lgrid(64, 48, 0.5 * np.pi / 180 / 60 * 40, 0.5 * np.pi / 180 / 60 * 40)
bin2d_case("fourier", m, np.arange(100, 3000, 40.0), rng, with_nan=True)
# 2. trim quirk: NO pixel above the last edge -> bincount is short and [1:-1] drops the last real bin
bin2d_case("trimquirk", m, np.linspace(0.0, m.max() * 1.5, 9), rng)
# 3. values exactly on edges (right=True), integer-valued radii, unsorted-looking real-space modrmap
yy, xx = np.mgrid[-20:21, -30:31]
r = np.sqrt(yy ** 2 + xx ** 2.0)
bin2d_case("onedge", r, np.array([0.0, 1.0, 2.0, 5.0, 10.0, 13.0, 25.0]), rng, with_nan=True)
# 4. non-uniform edges, first edge above the minimum, odd sizes
m = lgrid(37, 51, 1e-3, 1.3e-3)
bin2d_case("odd", m, np.array([200.0, 210.0, 500.0, 1000.0, 1100.0, 2500.0]), rng)

# Statistics: closed-form style inputs as orphics/tests/test_stats.py but stored numerically
for P in (1, 2, 3):
    parts = []
    allx = []
    for rank in range(P):
        s = rstats.Statistics(comm=None)
        rr = np.random.RandomState(100 + rank)
        xs = rr.standard_normal((rank + 2, 5)) + rank
        for x in xs:
            s.add("v", x)
        s.add_stack("m", (rank + 1) * np.ones((3, 4)) + rr.standard_normal((3, 4)))
        parts.append(s)
        allx.append(xs)
    # single-process reference of the cross-rank sum: feed everything to one Statistics
    tot = rstats.Statistics(comm=None)
    for rank in range(P):
        tot.extend("v", allx[rank])
        rr = np.random.RandomState(100 + rank)
        rr.standard_normal((rank + 2, 5))
        tot.add_stack("m", (rank + 1) * np.ones((3, 4)) + rr.standard_normal((3, 4)))
    tot.allreduce()
    np.savez_compressed(os.path.join(here, f"statistics_P{P}.npz"),
                        **{f"x{r}": allx[r] for r in range(P)},
                        N=tot.count("v"), mean=tot.mean("v"), cov=tot.cov("v"), var=tot.var("v"),
                        K=tot.stack_count("m"), stack=tot.stack_sum("m"))
    print("Statistics P", P, tot.count("v"))

cases = [(10, 3), (7, 7), (1024, 8), (13, 4), (5, 1), (1000, 6)]
np.savez_compressed(os.path.join(here, "mpi_distribute.npz"),
                    cases=np.array(cases),
                    **{f"n_{n}_{c}": rmpi.mpi_distribute(n, c)[0] for n, c in cases},
                    **{f"first_{n}_{c}": np.array([t[0] for t in rmpi.mpi_distribute(n, c)[1]]) for n, c in cases})
print("done")

# the reference's own checkpoint file (stats.py:1455-1484) for the P=2 case above
s = rstats.Statistics(comm=None)
z = np.load(os.path.join(here, "statistics_P2.npz"))
for r in range(2):
    s.extend("bandpowers", z[f"x{r}"])
s.add_stack("meanfield", np.arange(12.).reshape(3, 4))
s.allreduce()
s.save_reduced(os.path.join(here, "statistics_reduced_ref.npz"))
print("saved reference-format checkpoint")
