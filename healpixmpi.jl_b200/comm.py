"""Communicator plumbing.  The reference talks to MPI.jl (MPI.Comm, bcast, Gather(v), Alltoallv!);
here the same roles are played by torch.distributed process groups (gloo on CPU, nccl or gloo on
GPU boxes) behind a small MPI-flavoured facade, plus a single-process communicator.  The hot-path
exchange itself does NOT go through this module: the fused plan owns an NCCL communicator inside
libhpsht_b200.so; this facade only ships the NCCL unique id and serves Scatter!/Gather!."""
from __future__ import annotations

import numpy as np


class Comm:
    """Single-process communicator (MPI.COMM_SELF analogue)."""

    def rank(self) -> int:
        return 0

    def size(self) -> int:
        return 1

    def bcast(self, obj, root: int = 0):
        return obj

    def barrier(self) -> None:
        pass

    def gather(self, obj, root: int = 0):
        return [obj]

    def allgather(self, obj):
        return [obj]

    def alltoallv(self, send: np.ndarray, send_counts, recv: np.ndarray, recv_counts) -> None:
        assert send_counts[0] == recv_counts[0]
        recv[: recv_counts[0]] = send[: send_counts[0]]

    def allreduce_sum(self, value):
        return value

    def __eq__(self, other):
        return isinstance(other, Comm) and type(other) is type(self) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def _key(self):
        return ("self",)


COMM_SELF = Comm()


class TorchComm(Comm):
    """torch.distributed-backed communicator (MPI.COMM_WORLD analogue when group is None)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._dist = dist
        self.group = group

    def rank(self) -> int:
        return self._dist.get_rank(self.group)

    def size(self) -> int:
        return self._dist.get_world_size(self.group)

    def _key(self):
        return ("torch", id(self.group) if self.group is not None else 0)

    def bcast(self, obj, root: int = 0):
        box = [obj if self.rank() == root else None]
        self._dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def barrier(self) -> None:
        self._dist.barrier(group=self.group)

    def gather(self, obj, root: int = 0):
        out = [None] * self.size() if self.rank() == root else None
        self._dist.gather_object(obj, out, dst=root, group=self.group)
        return out

    def allgather(self, obj):
        out = [None] * self.size()
        self._dist.all_gather_object(out, obj, group=self.group)
        return out

    def alltoallv(self, send: np.ndarray, send_counts, recv: np.ndarray, recv_counts) -> None:
        """MPI.Alltoallv!(VBuffer(send, send_counts), VBuffer(recv, recv_counts), comm) on
        complex128 host buffers (src/sht.jl:32,133)."""
        import torch

        s = torch.from_numpy(np.ascontiguousarray(send).view(np.float64).reshape(-1, 2))
        r = torch.empty((int(sum(recv_counts)), 2), dtype=torch.float64)
        if self._dist.get_backend(self.group) == "nccl":
            s = s.cuda()
            r = r.cuda()
        self._dist.all_to_all_single(r, s, [int(c) for c in recv_counts],
                                     [int(c) for c in send_counts], group=self.group)
        recv[: int(sum(recv_counts))] = r.cpu().numpy().view(np.complex128).reshape(-1)

    def allreduce_sum(self, value):
        import torch

        t = torch.as_tensor(np.asarray(value, dtype=np.float64)).clone()
        if self._dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        self._dist.all_reduce(t, group=self.group)
        out = t.cpu().numpy()
        return out if out.ndim else float(out)


def comm_world() -> Comm:
    """COMM_WORLD: the default torch.distributed group if initialised, else the self communicator."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return TorchComm()
    except Exception:  # pragma: no cover
        pass
    return COMM_SELF
