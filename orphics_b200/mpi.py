"""orphics.mpi mirror for one 8xB200 box: the task split of mpi.py:78-102 with
torch.distributed (NCCL over NVLink; gloo on CPU) in place of mpi4py."""
import os

import numpy as np


def mpi_distribute(num_tasks, avail_cores, allow_empty=False):
    """Contiguous split of range(num_tasks); the remainder goes to the LAST ranks so that
    rank 0 never gets extra jobs (mpi.py:78-91)."""
    if not (allow_empty):
        assert avail_cores <= num_tasks
    min_each, rem = divmod(num_tasks, avail_cores)
    num_each = np.array([min_each] * avail_cores)
    if rem > 0:
        num_each[-rem:] += 1
    stops = np.cumsum(num_each).tolist()
    starts = [0] + stops[:-1]
    task_dist = [list(range(a, b)) for a, b in zip(starts, stops)]
    assert sum(num_each) == num_tasks
    return num_each, task_dist


class Comm:
    """The part of MPI.COMM_WORLD the hot path uses (Get_rank/Get_size/Barrier + an
    in-place sum all-reduce of device or host buffers)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.active = dist.is_available() and dist.is_initialized()

    def Get_rank(self):
        return self._dist.get_rank(self.group) if self.active else 0

    def Get_size(self):
        return self._dist.get_world_size(self.group) if self.active else 1

    def Barrier(self):
        if self.active:
            self._dist.barrier(self.group)

    def allreduce_sum_(self, tensor):
        """In-place SUM all-reduce of a torch tensor (CUDA tensor -> NCCL)."""
        if self.active and self.Get_size() > 1:
            self._dist.all_reduce(tensor, op=self._dist.ReduceOp.SUM, group=self.group)
        return tensor


def init_process_group(backend=None):
    """Join the torchrun world (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* from the environment);
    returns (rank, local_rank, world_size).  Without torchrun: (0, 0, 1), no group."""
    import torch
    import torch.distributed as dist
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1:
        return 0, 0, 1
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=ws, **kw)
    return rank, local, ws


def distribute(njobs, verbose=True, **kwargs):
    """mpi.py:95-102 over the torch.distributed world."""
    comm = Comm()
    rank, numcores = comm.Get_rank(), comm.Get_size()
    num_each, each_tasks = mpi_distribute(njobs, numcores, **kwargs)
    if rank == 0 and verbose:
        print("At most ", max(num_each), " tasks...")
    return comm, rank, each_tasks[rank]
