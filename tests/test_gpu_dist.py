"""GPU, 2 ranks (skipped with fewer than 2 GPUs): the exchange step of the path through the C ABI.  Two processes, one
per GPU, each run their shard; the packed Statistics triple [N | SUM | CROSS] and the packed mean-field stack
[stack | count] are summed with ONE ncclAllReduce each issued by liborphx.so (stats.py:1215-1217,1227-1228).  The
result must equal a single-GPU run of the whole job: N and the count exactly, sums to 1e-13."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    from orphics_b200 import _capi
    import ctypes as C
    n = C.c_int(0)
    _capi.lib.ox_device_count(C.byref(n))
    return n.value


def test_two_rank_nccl_allreduce_matches_single_gpu():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("NCCL_WORKER ")]
    assert out.returncode == 0 and lines, (out.stdout[-2000:], out.stderr[-3000:])
    rep = json.loads(lines[-1][len("NCCL_WORKER "):])
    assert rep["ok"] and rep["ws"] == 2 and rep["N"] == [12, 12] and rep["mf_count"] == [12, 12], rep


def test_single_rank_comm_is_a_noop():
    """nranks = 1: no NCCL needed, the reductions leave the accumulators untouched."""
    import numpy as np
    from orphics_b200 import _capi, mpi
    comm = mpi.NcclComm(0, 1)
    buf = _capi.DeviceBuffer(80).upload(np.arange(10.0))
    comm.allreduce_f64(buf.ptr, 10)
    assert np.array_equal(buf.download((10,), np.float64), np.arange(10.0))
    comm.free()
