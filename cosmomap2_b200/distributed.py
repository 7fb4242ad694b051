"""
Multi-GPU sharding of the map-making solve: one process per GPU, TOD partitioned by
(CES, detector) -- noise blocks and subscans never straddle detectors (reference layout,
interfaces/linearoperators.py:134-167, 609-615), so the time-domain stage needs no communication.

    A = sum_g P_g^T N_g^-1 P_g          one map-domain sum per A-matvec  (all_reduce, NCCL/NVLink)
    b = sum_g P_g^T N_g^-1 d_g          one sum at set-up
    moments (M_BD ingredients)          one sum at set-up, so every rank derives the same
                                        good-pixel mask / old2new and works on the same pixel set

PCG vectors are replicated: after the all-reduce every rank holds bit-identical q = A p, so the CG
dot products are computed redundantly and deterministically and need no communication at all.
``torch.distributed`` is the plumbing (backend "nccl" on GPUs; "gloo" is used by the CPU tests of
the host-side logic with NumPy-backed operators).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import linop as lp


def is_distributed(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def shard_detectors(ndet, world_size, rank):
    """Contiguous, balanced slice of detector indices for ``rank`` (first ranks get the extras)."""
    base, extra = divmod(int(ndet), int(world_size))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_tod(arrays, ndet, ns, world_size, rank):
    """Slice detector-major TOD arrays (length ndet*ns) down to this rank's detectors."""
    lo, hi = shard_detectors(ndet, world_size, rank)
    return [a[lo * ns:hi * ns] for a in arrays], (lo, hi)


def all_reduce_sum_(t, group=None):
    """In-place sum over ranks of a tensor (CUDA -> NCCL, CPU -> gloo)."""
    if is_distributed(group):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class AllReduceLO(lp.LinearOperator):
    """``sum_g A_g``: applies the local operator, then sums the map-domain result over ranks."""

    def __init__(self, local_op, group=None):
        self.local = local_op
        self.group = group
        super(AllReduceLO, self).__init__(local_op.nargin, local_op.nargout, matvec=self._run,
                                          symmetric=local_op.symmetric, device=True)

    def _run(self, x):
        y = self.local._apply(x)
        if y is x or y.data_ptr() == x.data_ptr():
            y = y.clone()
        return all_reduce_sum_(y, self.group)

    def _make_transpose(self):
        t = self.local.T
        if t is None:
            return None
        out = AllReduceLO(t, self.group)
        out._adjoint_of = self
        return out


class HostAllReduceLO(object):
    """NumPy twin of AllReduceLO for the gloo CPU tests of the sharding logic."""

    def __init__(self, local_matvec, n, group=None):
        self.local_matvec = local_matvec
        self.shape = (n, n)
        self.dtype = np.dtype(np.float64)
        self.group = group

    def matvec(self, x):
        y = torch.from_numpy(np.ascontiguousarray(self.local_matvec(np.asarray(x)), dtype=np.float64).copy())
        all_reduce_sum_(y, self.group)
        return y.numpy()

    def __mul__(self, x):
        return self.matvec(x)
