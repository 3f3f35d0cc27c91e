"""Communicator shim with the lowercase mpi4py API that Simulation and callbacks use (``sim.mpi.comm``).

The reference passes pickled Python objects through ``comm.scatter/bcast/reduce/gather/allgather/Barrier``
(simulation/simulation.py:275-282,363-365; callback/callback.py:95-136).  mpi4py is not part of this stack:
one process drives one GPU and ``torch.distributed`` is the transport, so this shim maps those calls onto
``torch.distributed`` object collectives (any backend; gloo on CPU, nccl on the GPU box).  Single-process runs
need no process group at all.
"""
from __future__ import annotations


class SingleComm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Barrier(self):
        pass

    def Dup(self):
        return self

    def bcast(self, obj, root=0):
        return obj

    def scatter(self, objs, root=0):
        return objs[0]

    def gather(self, obj, root=0):
        return [obj]

    def allgather(self, obj):
        return [obj]

    def reduce(self, obj, op=None, root=0):
        return obj

    def allreduce(self, obj, op=None):
        return obj

    def Abort(self, code=1):
        raise SystemExit(code)


class TorchComm:
    """torch.distributed-backed object collectives (process group must already be initialised)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group

    def Get_rank(self):
        return self.dist.get_rank(self.group)

    def Get_size(self):
        return self.dist.get_world_size(self.group)

    def Barrier(self):
        self.dist.barrier(self.group)

    def Dup(self):
        return self

    def bcast(self, obj, root=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def scatter(self, objs, root=0):
        out = [None]
        self.dist.scatter_object_list(out, objs if self.Get_rank() == root else None, src=root, group=self.group)
        return out[0]

    def allgather(self, obj):
        out = [None] * self.Get_size()
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def gather(self, obj, root=0):
        out = self.allgather(obj)
        return out if self.Get_rank() == root else None

    def allreduce(self, obj, op=None):
        vals = self.allgather(obj)
        fn = op if callable(op) else (lambda a, b: a + b)
        acc = vals[0]
        for v in vals[1:]:
            acc = fn(acc, v)
        return acc

    def reduce(self, obj, op=None, root=0):
        acc = self.allreduce(obj, op)
        return acc if self.Get_rank() == root else None

    def Abort(self, code=1):
        import os
        os._exit(code)


def default_comm():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return TorchComm()
    except Exception:
        pass
    return SingleComm()
