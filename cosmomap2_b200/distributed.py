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


class P2PAllReduce(object):
    """Sum of an n-vector over the ranks of ONE node through NVLink peer memory
    (cm2_allreduce_p2p: one kernel per rank, reduce-scatter by peer loads + broadcast by peer
    stores, flag barriers in peer memory).  torch only provides the plumbing: the buffers are torch
    CUDA allocations exported/imported with torch's CUDA-IPC reductions.

    Deterministic (fixed rank order) and bit-identical on every rank.  ``__call__(y)`` returns a
    view of the internal receive buffer, valid until the next call.
    """

    def __init__(self, n, group=None):
        import ctypes
        from torch.multiprocessing.reductions import reduce_tensor
        from . import _device as dv
        self.n = int(n)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("P2PAllReduce supports up to 8 ranks (one NVSwitch domain)")
        dev = dv.device()
        nsig = int(dv.call("cm2_allreduce_p2p_signal_bytes"))
        self.send = torch.empty(self.n + (self.n & 1), dtype=torch.float64, device=dev)
        self.recv = torch.empty(self.n + (self.n & 1), dtype=torch.float64, device=dev)
        self.sig = torch.zeros(nsig, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        payload = (dev.index, [reduce_tensor(t) for t in (self.send, self.recv, self.sig)])
        gathered = [None] * self.world
        dist.all_gather_object(gathered, payload, group=group)
        self._peers = []                       # keep the imported tensors alive
        send_p, recv_p, sig_p = [], [], []
        for g, (peer_dev, items) in enumerate(gathered):
            if g == self.rank:
                ts = (self.send, self.recv, self.sig)
            else:
                dv.call("cm2_enable_peer_access", int(peer_dev))
                # open every IPC mapping in THIS rank's device context (argument 6 of torch's
                # rebuild_cuda_tensor is the storage device): memory imported under the exporter's
                # device index is not reachable from kernels running on our device
                ts = tuple(fn(*(list(args[:6]) + [dev.index] + list(args[7:]))) for fn, args in items)
                self._peers.append(ts)
            send_p.append(ts[0].data_ptr())
            recv_p.append(ts[1].data_ptr())
            sig_p.append(ts[2].data_ptr())
        arr = ctypes.c_void_p * self.world
        self._send_tab, self._recv_tab, self._sig_tab = arr(*send_p), arr(*recv_p), arr(*sig_p)
        self.gen = 0
        torch.cuda.synchronize()

    def __call__(self, y):
        from . import _device as dv
        self.send[:self.n].copy_(y)
        self.gen += 1
        dv.call("cm2_allreduce_p2p", self._send_tab, self._recv_tab, self._sig_tab, self.rank, self.world,
                self.n, self.gen, dv.stream())
        return self.recv[:self.n]

    def error(self):
        """Non-zero if a flag wait timed out (a peer never arrived): results are then invalid."""
        torch.cuda.synchronize()
        return int(self.sig[-4:].view(torch.int32).item()) if self.sig.numel() % 4 == 0 else \
            int.from_bytes(bytes(self.sig[-4:].cpu().numpy().tobytes()), "little")

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self._peers = []


def p2p_enabled():
    import os
    return os.environ.get("CM2_P2P_ALLREDUCE", "1") != "0"


class AllReduceLO(lp.LinearOperator):
    """``sum_g A_g``: applies the local operator, then sums the map-domain result over ranks --
    through NVLink peer memory (P2PAllReduce) when the ranks share a node, else NCCL."""

    def __init__(self, local_op, group=None, p2p=None):
        self.local = local_op
        self.group = group
        self._p2p = None
        use = p2p_enabled() if p2p is None else p2p
        if (use and is_distributed(group) and torch.cuda.is_available() and dist.get_backend(group) == "nccl"
                and dist.get_world_size(group) <= 8):
            # the set-up can fail on ONE rank only (no peer access, IPC refused): all ranks then agree
            # to use NCCL, otherwise some would wait in the peer exchange for ranks that never come
            err = None
            try:
                self._p2p = P2PAllReduce(local_op.nargout, group)
            except Exception as e:
                err, self._p2p = e, None
            ok = torch.tensor([0.0 if self._p2p is None else 1.0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if ok.item() == 0.0:
                if err is not None or dist.get_rank(group) == 0:
                    import warnings
                    warnings.warn("P2P all-reduce unavailable (%s); all ranks use NCCL all_reduce" % (err,))
                self._p2p = None
        super(AllReduceLO, self).__init__(local_op.nargin, local_op.nargout, matvec=self._run,
                                          symmetric=local_op.symmetric, device=True)

    def check(self):
        """Raise if the peer-memory exchange ever timed out."""
        if self._p2p is not None and self._p2p.error() != 0:
            raise RuntimeError("P2P all-reduce: a peer did not arrive within the timeout (generation %d)"
                               % self._p2p.error())

    def close(self):
        if self._p2p is not None:
            self._p2p.close()
            self._p2p = None

    def _run(self, x):
        y = self.apply_transient(x)
        if self._p2p is not None:
            return y.clone()        # the exchange buffer is reused by the next call: hand out a copy
        return y

    def apply_transient(self, x):
        """Like ``_apply`` but the result may alias an internal buffer that the NEXT application
        overwrites (what the PCG loop wants: q = A p is consumed before A is applied again)."""
        y = self.local._apply(x)
        if self._p2p is not None:
            return self._p2p(y)
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


class ShardedTwoLevelPreconditionerLO(lp.LinearOperator):
    """``M_2lvl = M_BD (I - AZ E^-1 Z^T) + Z E^-1 Z^T`` with ``Z``, ``AZ`` and the ``M_BD`` blocks sharded
    by PIXEL over the ranks (each rank keeps 1/G of their rows: at configs[3] Z and AZ are the largest
    objects after the TOD), for PCG vectors that are replicated:

        t = Z_loc^T v_loc          tall-skinny kernel on the local rows
        t <- sum over ranks        r doubles (NCCL; bit-identical on every rank)
        c = E^-1 t                 replicated r x r apply
        y_loc = M_BD,loc (v_loc - AZ_loc c) + Z_loc c
        y <- all-gather(y_loc)     one n-vector, the same volume as the map exchange of an A apply

    Per apply every rank streams 3 r n / G doubles instead of 3 r n.  Built from the replicated
    operators (``Mbd``, ``DeflationLO(Z)``, ``DeflationLO(AZ)``, ``CoarseLO``): the local rows are copied,
    the caller may then drop the full ``Z`` / ``AZ``.
    """

    def __init__(self, Mbd, Zd, AZd, E, group=None):
        from . import _device as dv
        self.group = group
        dist_on = is_distributed(group)
        self.world = dist.get_world_size(group) if dist_on else 1
        self.rank = dist.get_rank(group) if dist_on else 0
        self.pol, npix, n = Mbd.pol, Mbd._n, Zd.nrows
        assert n == self.pol * npix and AZd.nrows == n and AZd.ncols == Zd.ncols
        self.r = int(Zd.ncols)
        spans = [shard_detectors(npix, self.world, g) for g in range(self.world)]
        plo, phi = spans[self.rank]
        self._lo, self._hi, self._npix_loc = self.pol * plo, self.pol * phi, phi - plo
        self._sizes = [self.pol * (b - a) for a, b in spans]
        self._nmax = max(self._sizes)
        self._z = Zd._zt[:, self._lo:self._hi].contiguous()
        self._az = AZd._zt[:, self._lo:self._hi].contiguous()
        self._inv = Mbd._inv_dev[6 * plo:6 * phi].clone()          # own, aligned copy of the local blocks
        self._einv = E._einv_dev
        nl = max(self._hi - self._lo, 1)
        self._v, self._u = dv.empty_f64(nl), dv.empty_f64(nl)
        self._y = dv.zeros_f64(self._nmax)
        self._t, self._c = dv.empty_f64(self.r), dv.empty_f64(self.r)
        self._gather = dv.empty_f64(self._nmax * self.world)
        self._work = dv.empty_f64(int(dv.call("cm2_defl_work_doubles", self.r)))
        super(ShardedTwoLevelPreconditionerLO, self).__init__(n, n, matvec=self.mult, symmetric=True, device=True)

    def mult(self, v):
        from . import _device as dv
        st = dv.stream
        nl, r = self._hi - self._lo, self.r
        self._v[:nl].copy_(v[self._lo:self._hi])
        dv.call("cm2_defl_zt_apply", dv.ptr(self._z), nl, r, nl, dv.ptr(self._v), 1, nl, dv.ptr(self._t),
                dv.ptr(self._work), st())
        all_reduce_sum_(self._t, self.group)
        dv.call("cm2_coarse_apply", dv.ptr(self._einv), r, dv.ptr(self._t), dv.ptr(self._c), st())
        dv.call("cm2_defl_z_apply", dv.ptr(self._az), nl, r, nl, dv.ptr(self._c), -1.0, 1.0, dv.ptr(self._v),
                dv.ptr(self._u), st())
        dv.call("cm2_bd_apply", dv.ptr(self._inv), self._npix_loc, self.pol, dv.ptr(self._u), dv.ptr(self._y), st())
        dv.call("cm2_defl_z_apply", dv.ptr(self._z), nl, r, nl, dv.ptr(self._c), 1.0, 1.0, dv.ptr(self._y),
                dv.ptr(self._y), st())
        if self.world == 1:
            return self._y[:nl].clone()
        dist.all_gather_into_tensor(self._gather, self._y, group=self.group)
        if all(s == self._nmax for s in self._sizes):
            return self._gather.clone()
        return torch.cat([self._gather[g * self._nmax:g * self._nmax + s] for g, s in enumerate(self._sizes)])
