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


def partition_pixels(npix, world_size):
    """Pixel slices [lo[g], lo[g+1]) of the ranks for the pixel-sharded solver: balanced, every interior
    boundary EVEN so that ``pol * lo`` doubles is 16-byte aligned for any ``pol``."""
    lo = [0]
    for g in range(1, int(world_size)):
        b = (int(npix) * g) // int(world_size)
        b -= b & 1
        lo.append(max(b, lo[-1]))
    lo.append(int(npix))
    return lo


def p2p_timeout_s():
    import os
    return float(os.environ.get("CM2_P2P_TIMEOUT", "20"))


class PeerBuffers(object):
    """Named CUDA buffers of this rank made visible to every rank of ONE node: torch allocations
    exported / imported with torch's CUDA-IPC reductions (plumbing only).  ``local[name]`` is this rank's
    tensor, ``table(name)`` a ctypes array of ``world`` device pointers (entry g = rank g's buffer mapped
    into this process), which is what the C ABI takes."""

    def __init__(self, tensors, group=None):
        import ctypes
        from torch.multiprocessing.reductions import reduce_tensor
        from . import _device as dv
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("peer-memory exchange supports up to 8 ranks (one NVSwitch domain)")
        dev = dv.device()
        self.names = list(tensors)
        self.local = dict(tensors)
        torch.cuda.synchronize()
        payload = (dev.index, [reduce_tensor(self.local[k]) for k in self.names])
        gathered = [None] * self.world
        dist.all_gather_object(gathered, payload, group=group)
        self._peers = []                       # keep the imported tensors alive
        ptrs = dict((k, []) for k in self.names)
        for g, (peer_dev, items) in enumerate(gathered):
            if g == self.rank:
                ts = [self.local[k] for k in self.names]
            else:
                dv.call("cm2_enable_peer_access", int(peer_dev))
                # open every IPC mapping in THIS rank's device context (argument 6 of torch's
                # rebuild_cuda_tensor is the storage device): memory imported under the exporter's
                # device index is not reachable from kernels running on our device
                ts = [fn(*(list(args[:6]) + [dev.index] + list(args[7:]))) for fn, args in items]
                self._peers.append(ts)
            for k, t in zip(self.names, ts):
                ptrs[k].append(t.data_ptr())
        arr = ctypes.c_void_p * self.world
        self._tables = dict((k, arr(*v)) for k, v in ptrs.items())
        torch.cuda.synchronize()

    def table(self, name):
        return self._tables[name]

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self._peers = []


class P2PAllReduce(object):
    """Sum of an n-vector over the ranks of ONE node through NVLink peer memory
    (cm2_allreduce_p2p: one kernel per rank, reduce-scatter by peer loads + broadcast by peer
    stores, flag barriers in peer memory).

    Deterministic (fixed rank order) and bit-identical on every rank.  ``__call__(y)`` returns a
    view of the internal receive buffer, valid until the next call; ``y`` may already BE the send
    buffer (``send_view``), in which case nothing is copied.
    """

    def __init__(self, n, group=None):
        from . import _device as dv
        self.n = int(n)
        self.group = group
        dev = dv.device()
        self.nsig = int(dv.call("cm2_allreduce_p2p_signal_bytes"))
        self.send = torch.empty(self.n + (self.n & 1), dtype=torch.float64, device=dev)
        self.recv = torch.empty(self.n + (self.n & 1), dtype=torch.float64, device=dev)
        self.sig = torch.zeros(self.nsig, dtype=torch.uint8, device=dev)
        self.buffers = PeerBuffers({"send": self.send, "recv": self.recv, "sig": self.sig}, group)
        self.world, self.rank = self.buffers.world, self.buffers.rank
        self.gen = 0
        dv.call("cm2_allreduce_p2p_set_timeout", p2p_timeout_s())

    @property
    def send_view(self):
        return self.send[:self.n]

    def __call__(self, y):
        from . import _device as dv
        dv.land(y, self.send_view)
        self.gen += 1
        dv.call("cm2_allreduce_p2p", self.buffers.table("send"), self.buffers.table("recv"), self.buffers.table("sig"),
                self.rank, self.world, self.n, self.gen, dv.stream())
        return self.recv[:self.n]

    def error(self):
        """Generation of the first flag wait that timed out (a peer never arrived), 0 if none; results
        from that generation on are invalid.  Synchronises the device."""
        torch.cuda.synchronize()
        return int(self.sig[self.nsig - 8:self.nsig - 4].view(torch.int32).item())

    def close(self):
        self.buffers.close()


def p2p_enabled():
    import os
    return os.environ.get("CM2_P2P_ALLREDUCE", "1") != "0"


def sharded_pcg_enabled():
    import os
    return os.environ.get("CM2_SHARDED_PCG", "1") != "0"


def agree_failed(flag, group=None):
    """True on every rank if ``flag`` is true on any rank (one tiny NCCL all-reduce)."""
    t = torch.tensor([1.0 if flag else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item() != 0.0)


class AllReduceLO(lp.LinearOperator):
    """``sum_g A_g``: applies the local operator, then sums the map-domain result over ranks --
    through NVLink peer memory (P2PAllReduce) when the ranks share a node, else NCCL.  The local
    operator writes straight into the peer-visible send buffer (no staging copy).

    ``cg(A, b, M=M_BD)`` does not use the all-reduce at all: it runs the pixel-sharded solver
    (ShardedPCG: reduce-scatter + M_BD + CG vector work + all-gather fused in one kernel per
    iteration) on the same exchange buffers."""

    def __init__(self, local_op, group=None, p2p=None):
        self.local = local_op
        self.group = group
        self._p2p = None
        self._sharded = {}
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
            if agree_failed(self._p2p is None, group):
                if err is not None or dist.get_rank(group) == 0:
                    import warnings
                    warnings.warn("P2P all-reduce unavailable (%s); all ranks use NCCL all_reduce" % (err,))
                self._p2p = None
        super(AllReduceLO, self).__init__(local_op.nargin, local_op.nargout, matvec=self._run,
                                          symmetric=local_op.symmetric, device=True)

    def check(self):
        """Raise if the peer-memory exchange ever timed out (local view; see ``recover``)."""
        if self._p2p is not None and self._p2p.error() != 0:
            raise RuntimeError("P2P all-reduce: a peer did not arrive within the timeout (generation %d)"
                               % self._p2p.error())

    def recover(self):
        """COLLECTIVE.  If the peer-memory exchange timed out on ANY rank, every rank drops it and uses
        NCCL from here on; returns True in that case (results since the failure are invalid: redo them)."""
        if self._p2p is None:
            return False
        if not agree_failed(self._p2p.error() != 0, self.group):
            return False
        self.disable_p2p("a peer-flag wait timed out")
        return True

    def disable_p2p(self, why=""):
        """COLLECTIVE: all ranks switch to the NCCL all-reduce (and the replicated PCG)."""
        if self._p2p is not None:
            import warnings
            if dist.get_rank(self.group) == 0:
                warnings.warn("peer-memory exchange disabled (%s): all ranks fall back to NCCL" % why)
            self._sharded = {}
            self._p2p.close()
            self._p2p = None

    def close(self):
        self._sharded = {}
        if self._p2p is not None:
            self._p2p.close()
            self._p2p = None

    def _run(self, x):
        y = self.apply_transient(x)
        if self._p2p is not None:
            return y.clone()        # the exchange buffer is reused by the next call: hand out a copy
        return y

    def apply_local_into(self, x, buf):
        """``buf <- A_local x`` without a staging copy when the local operator supports it."""
        from . import _device as dv
        with dv.map_output(buf):
            y = self.local._apply(x)
        return dv.land(y, buf)

    def apply_transient(self, x):
        """Like ``_apply`` but the result may alias an internal buffer that the NEXT application
        overwrites (what the PCG loop wants: q = A p is consumed before A is applied again)."""
        if self._p2p is not None:
            return self._p2p(self.apply_local_into(x, self._p2p.send_view))
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

    def sharded_solver(self, M):
        """The pixel-sharded PCG for ``cg(self, b, M=M)`` if it applies (peer memory available, M is the
        block-diagonal preconditioner of this map), else None.  Cached: the IPC set-up happens once."""
        from .linearoperators import BlockDiagonalPreconditionerLO
        if (self._p2p is None or not sharded_pcg_enabled() or not isinstance(M, BlockDiagonalPreconditionerLO)
                or M.size != self.nargin or self.nargin != self.nargout):
            return None
        hit = self._sharded.get(id(M))
        if hit is not None and hit.bd is M:
            return hit
        err = None
        try:
            solver = ShardedPCG(self, M)
        except Exception as e:
            err, solver = e, None
        if agree_failed(solver is None, self.group):
            if err is not None:
                import warnings
                warnings.warn("pixel-sharded PCG unavailable (%s)" % (err,))
            return None
        self._sharded[id(M)] = solver
        return solver


class ShardedPCG(object):
    """M_BD-preconditioned CG with the pixel-domain state sharded over the ranks of one node and the map
    exchange fused into the vector work (cm2_pcg_bd_sharded, csrc/pcg_sharded.cu): per iteration ONE
    local TOD pass (A_local p, written straight into the peer-visible buffer) and ONE kernel that
    reduce-scatters it by peer loads, does alpha / x / r / z = M_BD r / beta on the rank's pixel slice
    and all-gathers the next search direction by peer stores.  Same recurrence, exit rule and scalar
    workspace as ``pcg.PCG``; all ranks hold bit-identical scalars.  Rank g keeps 1/G of x, r, z, q and
    of the M_BD blocks; only p (and the local A p) are full length."""

    def __init__(self, A, Mbd):
        from . import _device as dv
        from .pcg import NSCAL
        self.A, self.bd, self.group = A, Mbd, A.group
        p2p = A._p2p
        self.world, self.rank = p2p.world, p2p.rank
        self.n, self.pol, self.npix = A.nargin, Mbd.pol, Mbd._n
        self.pix_lo = partition_pixels(self.npix, self.world)
        self.plo, self.phi = self.pix_lo[self.rank], self.pix_lo[self.rank + 1]
        self.elo, self.ehi = self.pol * self.plo, self.pol * self.phi
        nl = max(self.ehi - self.elo, 2)
        # x, r, z: full-length buffers of which this rank maintains its slice (elo is even: the views are 16-byte
        # aligned); a start from a replicated right-hand side fills them locally, without any exchange
        self._xf, self._rf, self._zf = (dv.zeros_f64(max(self.n, 2) + 2) for _ in range(3))
        self._xs, self._rs, self._zs = (t[self.elo:self.elo + nl] for t in (self._xf, self._rf, self._zf))
        self._qs = dv.zeros_f64(nl)
        self._inv = Mbd._inv_dev[6 * self.plo:6 * self.phi].clone() if self.phi > self.plo else dv.zeros_f64(6)
        self.scal = dv.zeros_f64(NSCAL)
        self._part = dv.empty_f64(int(dv.call("cm2_pcg_sharded_work_doubles")))
        self._sig = torch.zeros(int(dv.call("cm2_pcg_sharded_signal_bytes")), dtype=torch.uint8, device=dv.device())
        self._sigbuf = PeerBuffers({"sig": self._sig}, self.group)
        self._y, self._p = p2p.send_view, p2p.recv[:self.n]
        self._tabs = self._tables()
        self.gen = 0
        self._pin = [torch.empty(NSCAL, dtype=torch.float64).pin_memory() for _ in range(2)]
        self._ev = [torch.cuda.Event(), torch.cuda.Event()]
        self._queued = 0
        self._snap = 0
        self._b = None

    def _tables(self):
        import ctypes
        arr = ctypes.c_void_p * self.world
        i64 = ctypes.c_int64 * (self.world + 1)

        def own(t):
            a = arr()
            a[self.rank] = t.data_ptr()
            return a
        return dict(pix_lo=i64(*self.pix_lo), x=own(self._xs), r=own(self._rs), z=own(self._zs), q=own(self._qs),
                    inv=own(self._inv), scal=own(self.scal), part=own(self._part), arr=arr)

    # ---- slices ---------------------------------------------------------------------------------
    @property
    def x(self):
        """This rank's slice of the solution (elements [elo, ehi) of the map vector)."""
        return self._xs[:self.ehi - self.elo]

    def slice_of(self, v):
        return v[self.elo:self.ehi]

    def _launch(self, reset, b_tab, atol, rtol):
        from . import _device as dv
        t, p2p = self._tabs, self.A._p2p
        self.gen += 1
        dv.call("cm2_pcg_bd_sharded", 1 if reset else 0, self.pol, self.world, self.rank, 1, t["pix_lo"],
                p2p.buffers.table("send"), p2p.buffers.table("recv"), self._sigbuf.table("sig"), t["x"], t["r"], t["z"],
                t["q"], t["inv"], b_tab, t["scal"], t["part"], self.gen, float(atol), float(rtol), p2p_timeout_s(),
                dv.stream())

    # ---- the PCG state machine (same surface as pcg.PCG) ---------------------------------------------
    def start(self, b, x0=None, atol=0.0, rtol=0.0):
        """x <- 0, r <- b, z = M r, p = z on every rank, rho, ||r||^2, atol_eff = max(atol, rtol ||b||).
        ``b``: the full right-hand side (CUDA, replicated) or this rank's slice of it."""
        from . import _device as dv
        if x0 is not None:
            raise ValueError("ShardedPCG starts from x0 = 0")
        if b.numel() == self.n and self.n != self.ehi - self.elo:
            # replicated right-hand side: every rank computes r = b, z = M_BD b, the full first search direction
            # p = z and the scalars by itself (same data, same kernel, same order: bit-identical on all ranks) --
            # no all-gather, no flag round.  Safe without a barrier: the last kernel of this rank on the stream
            # ended with the end barrier of its iteration, after which no peer writes this rank's p or reads its y.
            self._b = b
            dv.call("cm2_pcg_bd_reset", dv.ptr(self.bd._inv_dev), self.npix, self.pol, dv.ptr(self._rf), dv.ptr(self._zf),
                    dv.ptr(self.scal), float(atol), float(rtol), dv.ptr(b), dv.ptr(self._xf), dv.ptr(self._p), dv.stream())
            self._queued = 0
            return
        bs = b if b.numel() == self.ehi - self.elo and self.world > 1 else self.slice_of(b)
        if bs.numel() == 0:
            bs = self._zs                      # an empty slice: any valid pointer
        self._b = bs                           # keep alive until the kernel has run
        bt = self._tabs["arr"]()
        bt[self.rank] = bs.data_ptr()
        self._launch(True, bt, atol, rtol)
        self._queued = 0

    def step_async(self):
        """Queue one iteration: the local TOD pass into the peer-visible buffer, then the fused kernel."""
        self.A.apply_local_into(self._p, self._y)
        self._launch(False, None, 0.0, 0.0)
        self._queued += 1

    def state(self):
        s = self.scal.cpu()
        return float(np.sqrt(s[3].item())), bool(s[7].item() != 0.0), int(s[8].item())

    def _snapshot(self):
        k = self._snap % 2
        self._pin[k].copy_(self.scal, non_blocking=True)
        self._ev[k].record()
        self._snap += 1

    def _read_snapshot(self, idx):
        k = idx % 2
        self._ev[k].synchronize()
        s = self._pin[k]
        return float(np.sqrt(s[3].item())), bool(s[7].item() != 0.0), int(s[8].item())

    def tick(self):
        self._snapshot()
        if self._snap >= 2:
            return self._read_snapshot(self._snap - 2)
        return None

    def failed(self):
        """Local view: non-zero if one of this rank's flag waits timed out (synchronises)."""
        return int(self.scal[9].item()) != 0

    def gather_x(self):
        """The full solution on every rank (one all-gather of the slices, outside the iteration)."""
        sizes = [self.pol * (self.pix_lo[g + 1] - self.pix_lo[g]) for g in range(self.world)]
        nmax = max(max(sizes), 1)
        mine = torch.zeros(nmax, dtype=torch.float64, device=self._xs.device)
        mine[:sizes[self.rank]].copy_(self.x)
        full = torch.empty(nmax * self.world, dtype=torch.float64, device=self._xs.device)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        if all(s == nmax for s in sizes):
            return full
        return torch.cat([full[g * nmax:g * nmax + s] for g, s in enumerate(sizes)])


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
