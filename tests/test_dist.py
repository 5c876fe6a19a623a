"""CPU tests of the N>1 path: world_size-2 gloo process group, the reference's
closed-form-in-P Statistics tests (orphics/tests/test_stats.py:12-183) re-expressed over
torch.distributed, and the mpi_distribute sharding rule."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys, numpy as np
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from orphics_b200 import mpi, stats
    rank, local, ws = mpi.init_process_group("gloo")
    assert ws == 2
    comm, r, my = mpi.distribute(11, verbose=False)
    assert r == rank and my == ([0,1,2,3,4] if rank == 0 else [5,6,7,8,9,10])
    # test_stats.py closed forms: rank r adds r+1 samples of value (r+1)*[1,2,3]
    s = stats.Statistics(comm=True)
    for _ in range(rank + 1):
        s.add("v", (rank + 1) * np.array([1., 2., 3.]))
    s.add_stack("m", (rank + 1) * np.ones((2, 2)))
    if rank == 1:
        s.add("only1", np.array([5., 7.]))          # a label present on a subset of ranks
    s.allreduce()
    P = ws
    N = P * (P + 1) // 2
    assert s.count("v") == N
    tot = sum((q + 1) * (q + 1) for q in range(P))
    np.testing.assert_allclose(s.mean("v"), np.array([1., 2., 3.]) * tot / N, rtol=0, atol=0)
    sq = sum((q + 1) ** 3 for q in range(P))
    var = (sq - tot * tot / N) / (N - 1)
    np.testing.assert_allclose(s.cov("v"), var * np.outer([1, 2, 3], [1, 2, 3]), rtol=1e-14)
    np.testing.assert_allclose(s.var("v"), np.diag(s.cov("v")), rtol=1e-14)
    np.testing.assert_allclose(s.stack_sum("m"), np.ones((2, 2)) * P * (P + 1) / 2, rtol=0, atol=0)
    assert s.stack_count("m") == P
    assert s.count("only1") == 1
    np.testing.assert_array_equal(s.mean("only1"), [5., 7.])
    # bandpower-triple merge as bench.py does it across ranks
    t = stats.Statistics(comm=True)
    x = np.arange(6.).reshape(2, 3) + 10 * rank
    t.add_triple("bp", 2, x.sum(0), x.T @ x)
    t.allreduce()
    allx = np.concatenate([np.arange(6.).reshape(2, 3) + 10 * q for q in range(P)])
    np.testing.assert_allclose(t.mean("bp"), allx.mean(0), rtol=1e-14)
    np.testing.assert_allclose(t.cov("bp"), np.cov(allx.T), rtol=1e-12)
    dist.destroy_process_group()
    print("RANK_OK", rank)
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gloo_world2_statistics_and_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"RANK_OK {rank}" in o


def test_single_process_statistics_matches_reference_golden():
    from conftest import load_golden
    from orphics_b200 import stats
    for P in (1, 2, 3):
        z = load_golden(f"statistics_P{P}.npz")
        s = stats.Statistics()
        for r in range(P):
            s.extend("v", z[f"x{r}"])
        s.allreduce()
        assert s.count("v") == int(z["N"])
        np.testing.assert_allclose(s.mean("v"), z["mean"], rtol=1e-14)
        np.testing.assert_allclose(s.cov("v"), z["cov"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(s.var("v"), z["var"], rtol=1e-12, atol=1e-14)


def test_statistics_checkpoint_format_matches_reference(tmp_path):
    """save_reduced/load_reduced use the reference's .npz key layout (stats.py:1455-1530):
    a file written by the reference loads here, and our file has the same keys and arrays."""
    from conftest import GOLDEN, load_golden
    from orphics_b200 import stats
    ref = np.load(os.path.join(GOLDEN, "statistics_reduced_ref.npz"))
    s = stats.Statistics.load_reduced(os.path.join(GOLDEN, "statistics_reduced_ref.npz"))
    z = load_golden("statistics_P2.npz")
    assert s.count("bandpowers") == int(z["N"])
    np.testing.assert_allclose(s.mean("bandpowers"), z["mean"], rtol=1e-14)
    np.testing.assert_allclose(s.cov("bandpowers"), z["cov"], rtol=1e-12, atol=1e-14)
    np.testing.assert_array_equal(s.stack_sum("meanfield"), np.arange(12.).reshape(3, 4))
    mine = stats.Statistics()
    for r in range(2):
        mine.extend("bandpowers", z[f"x{r}"])
    mine.add_stack("meanfield", np.arange(12.).reshape(3, 4))
    mine.allreduce()
    out = tmp_path / "ours.npz"
    mine.save_reduced(out)
    ours = np.load(out)
    assert sorted(ours.files) == sorted(ref.files)
    for k in ref.files:
        assert ours[k].dtype == ref[k].dtype and ours[k].shape == ref[k].shape
        np.testing.assert_allclose(ours[k], ref[k], rtol=1e-13)
