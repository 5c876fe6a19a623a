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


class NcclComm:
    """Communicator of the C ABI (ox_comm_*, include/orphx.h): the data plane of the one exchange step of the
    path -- Statistics.allreduce (stats.py:1184-1232) -- as ncclAllReduce(sum) over NVLink issued by
    liborphx.so on its own stream.  torch.distributed (or any other rendezvous) is only used to hand rank 0's
    128-byte NCCL unique id to the other ranks.  With one rank no NCCL is needed and reductions are no-ops."""

    def __init__(self, rank=0, nranks=1, group=None, unique_id=None):
        import ctypes as C
        from ._capi import lib, check, ptr, require_device
        require_device()
        self.rank, self.nranks = int(rank), int(nranks)
        uid = np.zeros(128, dtype=np.uint8)
        if self.nranks > 1:
            if unique_id is not None:
                uid[:] = np.frombuffer(bytes(unique_id), dtype=np.uint8)[:128]
            else:
                import torch.distributed as dist
                if self.rank == 0:
                    check(lib.ox_comm_unique_id(ptr(uid), uid.nbytes))
                box = [uid.tobytes()]
                dist.broadcast_object_list(box, src=0, group=group)
                uid = np.frombuffer(box[0], dtype=np.uint8).copy()
        h = C.c_void_p()
        check(lib.ox_comm_create(self.rank, self.nranks, ptr(uid), uid.nbytes, C.byref(h)))
        self.handle = h

    @staticmethod
    def new_unique_id():
        from ._capi import lib, check, ptr
        uid = np.zeros(128, dtype=np.uint8)
        check(lib.ox_comm_unique_id(ptr(uid), uid.nbytes))
        return uid.tobytes()

    def nccl_version(self):
        import ctypes as C
        from ._capi import lib, check
        r, n, v = C.c_int(), C.c_int(), C.c_int()
        check(lib.ox_comm_info(self.handle, C.byref(r), C.byref(n), C.byref(v)))
        return v.value

    def allreduce_f64(self, dev_ptr, count):
        """In-place sum of count float64 values at a device address."""
        import ctypes as C
        from ._capi import lib, check
        check(lib.ox_comm_allreduce_f64(self.handle, C.c_void_p(int(dev_ptr)), int(count)))

    def free(self):
        from ._capi import lib
        if getattr(self, "handle", None):
            lib.ox_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
