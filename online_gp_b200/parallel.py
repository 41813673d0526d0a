"""Row-sharded multi-GPU WISKI: the inducing grid is partitioned across ranks along its slowest axis.

One process per GPU (``torch.distributed``; NCCL over NVLink on B200, gloo in the CPU tests).  Rank ``g`` of ``G``
owns the contiguous slab of rows ``[g m/G, (g+1) m/G)`` (``g_0 / G`` hyper-planes of grid axis 0) of the root panel
``L``, the inverse-root panel ``B``, ``K L`` and ``interpolation_cache``.  Replicated on every rank: the
hyper-parameters, the Toeplitz columns and every r x r object (``Q``, its Cholesky factor, the rank-q update
factors), so the small dense algebra and its autograd run identically everywhere.

Exchange steps per streaming step
  * projection ``p = B^T v`` (r x q), ``c = L^T K b`` (r), predictive gathers: all-reduce of partial sums;
  * ``Q - I = L^T (K L)``: all-reduce of the r x r Gram partials;
  * ``K L``: grid axes 1..d-1 are slab-local; the axis-0 Toeplitz factor mixes slabs and is applied in a
    *column-sharded* layout reached by one all-to-all each way (m r b / G bytes per rank each);
  * hyper-gradient (column gradient of the Kronecker operator): the same exchange for the prefix chain and for the
    axis-0 contraction, then one all-reduce of the d x g column gradient;
  * ``K b`` for the m-vector ``b``: all-gather of b (m b bytes), replicated MVM;
  * CG path (r > max_cholesky_size): one all-reduce of r x c per iteration (``sharded_cg_solve``).
Every autograd Function below returns *complete* (replicated) gradients for replicated inputs, so hyper-parameter
updates need no extra synchronisation.

The all-to-all (fused 32^4 path): ``K L`` and its gradient live as ``world`` column blocks ``[world, m_loc, r / world]``
— the receive buffer of one exchange and the send buffer of the next — which the pair kernels and the tensor-core
Gram / panel GEMM read and write in place (no transposing copies).  When symmetric memory is available the send
buffer is peer-mapped and every rank *pulls* its chunks over NVLink (``Comm.peer_exchange``); otherwise NCCL's
``all_to_all_single`` moves the same buffers.  ``enable_cuda_graphs()`` replays the whole step, collectives included,
from two captured CUDA graphs.

Mirrors, for one output and an Identity stem, ``OnlineSKIRegression.evaluate`` / ``.update``
(``online_gp/models/online_ski_regression.py:64-78,113-146``) on top of the same kernels as the single-GPU path.
"""
import ctypes
import math
import os
import warnings

import torch
import torch.distributed as dist

from . import ops, settings
from .graphs import StepGraphs, make_adam_capturable
from .kernels import GridInterpolationKernel, RBFKernel, ScaleKernel
from .lazy.updated_root_lazy_tensor import _sym_factors
from .likelihoods import FNMGLikelihood


# ---------------------------------------------------------------------------------------------- communication
class _PushBuffers:
    """Four peer-mapped regions (A..D, `numel` elements each) of one symmetric allocation per rank: the receive buffers
    of the kernels that push their result straight into the consumers' memory over NVLink (``ops.*_push``).

    Use in one streaming step (W ranks, rank p, panel [m_loc, c], cw = c / W, part = m_loc * cw elements):
      A  (T_a x T_b) L        pushed by the slab-local pair pass      -> all rows of my columns   [W m_loc, cw]
      B  K L                  pushed by the column-sharded pair pass  -> my rows of every block   [W, m_loc, cw]
      C  L grad_Q             pushed by the panel GEMM's epilogue      -> all rows of my columns
      D  (T_0 x T_v) Z        pushed by the column-sharded grad pass   -> my rows of every block
    Rank p always writes part p of a region on rank j.  Every push is followed by ``barrier()`` (all ranks' producers
    have finished) before the region is read; a region is rewritten one step later, with at least one barrier after
    its last reader on every rank (A, C: read by the pass that pushes D; B: by the Gram / prediction ops that precede the
    push of C; D: by the last pair pass, followed by next step's barrier after A)."""

    REGIONS = ("A", "B", "C", "D")

    def __init__(self, buf, hdl, peers, numel, rank, world):
        self.buf, self.hdl, self.peers = buf, hdl, peers
        self.numel, self.rank, self.world = numel, rank, world
        self._tables = {}
        self.c_pushed = False        # region C holds the pushed gradient panel of the Gram backward (consumed once)

    def local(self, name, numel=None):
        k = self.REGIONS.index(name)
        return self.buf[k * self.numel:k * self.numel + (self.numel if numel is None else numel)]

    def dst(self, name, part):
        """ctypes table: for every rank j the device address of part `rank` (part elements) of region `name` on rank j."""
        key = (name, part)
        if key not in self._tables:
            k = self.REGIONS.index(name)
            isz = self.buf.element_size()
            self._tables[key] = (ctypes.c_void_p * self.world)(
                *[pv.data_ptr() + (k * self.numel + self.rank * part) * isz for pv in self.peers])
        return self._tables[key]

    def barrier(self):
        self.hdl.barrier(0)


class Comm:
    """Thin wrapper over a torch.distributed process group (None => single process, every collective a no-op)."""

    def __init__(self, group=None):
        self.enabled = dist.is_available() and dist.is_initialized()
        self.group = group
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        self._a2a_ok = self.enabled and dist.get_backend(group) != "gloo"
        self.xbuf = self._xhdl = self._xpeers = self._xstreams = None      # peer-memory exchange (enable_peer_exchange)
        self.push = None                                                    # _PushBuffers (enable_push)
        self._far = None                                                    # (symmetric buffer, group name): enable_fast_allreduce

    def enable_push(self, numel, dtype, device):
        """Collective.  Sets up the peer-mapped receive regions of the pushing kernels (``_PushBuffers``); returns them,
        or None when peer memory is unavailable or WISKI_PUSH_EXCHANGE=0 (then the exchanges run as separate passes)."""
        self.push = None
        if (self.world == 1 or self.world > 8 or not self._a2a_ok or dtype != torch.float32
                or os.environ.get("WISKI_PUSH_EXCHANGE", "1") == "0"):
            return None
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else dist.group.WORLD
            gname = group.group_name
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    symm.enable_symm_mem_for_group(gname)
            except Exception:                       # noqa: BLE001 - newer torch enables groups implicitly
                pass
            buf = symm.empty(4 * numel, dtype=dtype, device=device)
            hdl = symm.rendezvous(buf, gname)
            peers = [hdl.get_buffer(p, (4 * numel,), dtype) for p in range(self.world)]
            if any(pv.device != buf.device for pv in peers):
                ok = 0
        except Exception as err:                    # noqa: BLE001
            warnings.warn(f"peer-memory push unavailable ({type(err).__name__}: {err})")
            ok = 0
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            return None
        buf.zero_()
        self.push = _PushBuffers(buf, hdl, peers, numel, self.rank, self.world)
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)
        return self.push

    def push_buffers(self, numel, like):
        pb = self.push
        if pb is None or pb.buf.dtype != like.dtype or pb.numel != numel:
            return None
        return pb

    def allreduce_(self, t):
        if self.world > 1:
            far = self._far
            if (far is not None and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                    and 0 < t.numel() <= far[0].numel() and (t.numel() * 4) % 16 == 0):
                # latency-sized sums (r x r Gram partials, r x q projections, gathered stencil rows): every rank copies its
                # addend into its symmetric buffer and reads all peers' over NVLink in ONE kernel — no ring, no protocol
                # hand-shakes (NCCL's LL all-reduce takes several times longer for these sizes on 8 GPUs)
                res = torch.ops.symm_mem.one_shot_all_reduce_copy(far[0][:t.numel()], t.view(-1), "sum", far[1])
                t.view(-1).copy_(res)
                return t
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def enable_fast_allreduce(self, device, max_elems=1 << 19):
        """Collective.  One symmetric fp32 buffer (2 MB) for the one-shot all-reduce of small replicated quantities;
        WISKI_SYMM_ALLREDUCE=0 keeps every sum on NCCL."""
        self._far = None
        if self.world == 1 or not self._a2a_ok or os.environ.get("WISKI_SYMM_ALLREDUCE", "1") == "0":
            return False
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else dist.group.WORLD
            gname = group.group_name
            buf = symm.empty(max_elems, dtype=torch.float32, device=device)
            symm.rendezvous(buf, gname)
            probe = torch.ones(8, dtype=torch.float32, device=device)
            res = torch.ops.symm_mem.one_shot_all_reduce_copy(buf[:8], probe, "sum", gname)
            if float(res.sum().item()) != 8.0 * self.world:
                ok = 0
        except Exception as err:                    # noqa: BLE001
            warnings.warn(f"symmetric-memory all-reduce unavailable ({type(err).__name__}: {err}); using NCCL")
            ok = 0
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            return False
        self._far = (buf, gname)
        return True

    def allgather(self, t):
        """[world, *t.shape]"""
        if self.world == 1:
            return t.unsqueeze(0)
        out = torch.empty(self.world, *t.shape, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group) if self._a2a_ok else \
            dist.all_gather(list(out.unbind(0)), t.contiguous(), group=self.group)
        return out

    # ---- peer-memory exchange (NVLink / NVSwitch): the all-to-all of the row <-> column layout change as direct
    # pulls from the peers' send buffers instead of NCCL's send/recv kernel (measured on 2 x B200: 208 GB/s per
    # direction for NCCL's all_to_all of 470 MB, far below what the copy engines reach over NVLink 5).
    # The send buffer is ONE symmetric allocation per model, mapped into every rank's address space
    # (torch.distributed._symmetric_memory); producers (fused pair kernels, panel GEMM) write straight into it.
    def enable_peer_exchange(self, numel, dtype, device):
        """Collective.  Returns the local symmetric send buffer (flat, `numel` elements) or None when peer memory is
        unavailable (then ``all_to_all`` goes through NCCL).  Every rank gets the same answer."""
        self.xbuf = self._xhdl = self._xpeers = self._xstreams = None
        if self.world == 1 or not self._a2a_ok or os.environ.get("WISKI_PEER_EXCHANGE", "1") == "0":
            return None
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else dist.group.WORLD
            gname = group.group_name
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    symm.enable_symm_mem_for_group(gname)
            except Exception:                       # noqa: BLE001 - newer torch enables groups implicitly
                pass
            buf = symm.empty(numel, dtype=dtype, device=device)
            hdl = symm.rendezvous(buf, gname)
            peers = [hdl.get_buffer(p, (numel,), dtype) for p in range(self.world)]
            if any(pv.device != buf.device for pv in peers):
                ok = 0
        except Exception as err:                    # noqa: BLE001
            warnings.warn(f"peer-memory exchange unavailable ({type(err).__name__}: {err}); using NCCL all_to_all")
            ok = 0
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            return None
        self.xbuf, self._xhdl, self._xpeers = buf, hdl, peers
        return buf

    def send_buffer(self, numel, like):
        """The symmetric send buffer if it is set up and fits (numel elements of like's dtype), else None."""
        xb = self.xbuf
        if xb is None or xb.dtype != like.dtype or xb.numel() < numel:
            return None
        return xb[:numel]

    def peer_exchange(self, shape):
        """All-to-all of the symmetric send buffer viewed as ``shape`` = [world, rows, cw] (chunk j goes to rank j):
        barrier (every rank's producer has finished), pull my chunk from every peer, barrier (buffer reusable)."""
        W, rank = self.world, self.rank
        self._xhdl.barrier(0)
        n = shape[0] * shape[1] * shape[2]
        recv = torch.empty(shape, dtype=self.xbuf.dtype, device=self.xbuf.device)
        # one copy engine does not fill NVLink (measured: 367 GB/s for a single 235 MB pull), so the pulls are cut
        # into pieces issued on side streams (fork / join with events: also valid under CUDA-graph capture)
        cur = torch.cuda.current_stream()
        if self._xstreams is None:
            self._xstreams = [torch.cuda.Stream(device=self.xbuf.device) for _ in range(8)]
        pieces = []
        nsplit = max(1, 8 // max(1, W - 1))
        for k in range(1, W):
            p = (rank + k) % W                      # stagger the peers
            src, dst = self._xpeers[p][:n].view(shape)[rank].reshape(-1), recv[p].reshape(-1)
            step = -(-src.numel() // nsplit)
            pieces += [(dst[o:o + step], src[o:o + step]) for o in range(0, src.numel(), step)]
        fork = torch.cuda.Event()
        fork.record(cur)
        joins = []
        for i, (dst, src) in enumerate(pieces):
            st = self._xstreams[i % len(self._xstreams)]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                dst.copy_(src)
                ev = torch.cuda.Event()
                ev.record(st)
            joins.append(ev)
        recv[rank].copy_(self._xpeers[rank][:n].view(shape)[rank])          # local chunk on the main stream
        for ev in joins:
            cur.wait_event(ev)
        self._xhdl.barrier(0)
        return recv

    def all_to_all(self, send):
        """send [world, ...] (chunk j goes to rank j) -> recv [world, ...] (chunk i came from rank i)."""
        if self.world == 1:
            return send
        if self.xbuf is not None and send.untyped_storage().data_ptr() == self.xbuf.untyped_storage().data_ptr():
            return self.peer_exchange(tuple(send.shape))
        send = send.contiguous()
        if self._a2a_ok:
            recv = torch.empty_like(send)
            dist.all_to_all_single(recv, send, group=self.group)
            return recv
        # gloo (CPU tests) has no all_to_all: emulate with an all-gather
        return self.allgather(send)[:, self.rank].contiguous()


class ShardPlan:
    """Partition of the m = prod(sizes) grid rows along grid axis 0."""

    def __init__(self, sizes, world, rank):
        self.sizes = [int(s) for s in sizes]
        self.world, self.rank = world, rank
        if self.sizes[0] % world != 0:
            raise ValueError(f"grid axis 0 ({self.sizes[0]} points) must be divisible by the number of ranks ({world})")
        self.d = len(self.sizes)
        self.m = 1
        for s in self.sizes:
            self.m *= s
        self.g0_loc = self.sizes[0] // world
        self.rest = self.m // self.sizes[0]
        self.m_loc = self.g0_loc * self.rest
        self.row0 = rank * self.m_loc

    def localize(self, idx, val):
        """Global stencils -> slab-local stencils (entries owned by other ranks get weight 0)."""
        mask = (idx >= self.row0) & (idx < self.row0 + self.m_loc)
        return torch.where(mask, idx - self.row0, torch.zeros_like(idx)), torch.where(mask, val, torch.zeros_like(val))

    def local_axes(self, c):
        """(axis, g, outer, inner) for the slab-local grid axes 1..d-1 of a [m_loc, c] panel."""
        out = []
        outer = self.g0_loc
        for i in range(1, self.d):
            inner = (self.m_loc // (outer * self.sizes[i])) * c
            out.append((i, self.sizes[i], outer, inner))
            outer *= self.sizes[i]
        return out


def _to_cols(X, plan, comm):
    """[m_loc, c] row slab -> [g0 * rest, c / world] column block (all rows of my columns)."""
    W = plan.world
    if W == 1:
        return X
    m_loc, c = X.shape
    send = X.view(m_loc, W, c // W).permute(1, 0, 2)
    return comm.all_to_all(send).reshape(W * m_loc, c // W)


def _to_rows(Xc, plan, comm):
    """inverse of _to_cols."""
    W = plan.world
    if W == 1:
        return Xc
    cw = Xc.shape[1]
    recv = comm.all_to_all(Xc.view(W, plan.m_loc, cw))
    return recv.permute(1, 0, 2).reshape(plan.m_loc, W * cw)


def _apply_local_axes(X, cols, plan):
    for i, g, outer, inner in plan.local_axes(X.shape[1]):
        X = ops.kron_axis_apply(X, cols[i], g, outer, inner)
    return X


def _fused_ok(plan, X):
    """Fused two-axes-per-pass kernels usable: fp32 on CUDA, d = 4 with 32-point axes 1..3 and a 32-point axis 0,
    and 16-column tiles both in the row-sharded (c) and the column-sharded (c / world) layout."""
    return (X.is_cuda and X.dtype == torch.float32 and plan.d == 4 and all(s == 32 for s in plan.sizes)
            and X.shape[1] % (16 * plan.world) == 0)


class _ShardedKronFn(torch.autograd.Function):
    """(K X)_loc for a row-sharded panel X (c divisible by world); complete gradient w.r.t. cols.

    The result is returned as column *blocks* ``[nb, m_loc, c / nb]`` (block j = columns [j c/nb, (j+1) c/nb)):
    on the fused path nb = world and the blocks are exactly the receive buffer of the column -> row all-to-all, so
    neither side of either exchange needs a transposing copy (the pair kernels read / write the chunked layout
    directly, ``wiski_kron_fused_pair_*_lay_f32``); otherwise nb = 1."""

    @staticmethod
    def forward(ctx, cols, X, plan, comm, dirs=None):
        cols = cols.contiguous()
        ctx.plan, ctx.comm = plan, comm
        ctx.fused = _fused_ok(plan, X)
        ctx.dirs = None if (dirs is None or not ctx.fused) else dirs.detach().to(cols.dtype).contiguous()
        W = plan.world
        if ctx.fused:
            # slab [g0/W, 32, 32, 32, c]: one pair of the axes 1..3 is slab-local, the pair containing axis 0 runs in the
            # column-sharded layout: (1,2) + (0,3) with the directional tensor-core passes, (2,3) + (0,1) otherwise
            ctx.col_pair, ctx.loc_pair = ops.kron_pairs(plan.sizes, directional=ctx.dirs is not None)
            slab = [plan.g0_loc] + plan.sizes[1:]
            cw = X.shape[1] // W
            pb = comm.push_buffers(X.numel(), X) if (ctx.dirs is not None and W > 1) else None
            ctx.pushed = pb is not None
            if pb is not None:
                # compute + exchange in one kernel each: the pair passes store into the peers' regions A / B
                part = plan.m_loc * cw
                ops._fused_pair_apply_push(cols, slab, ctx.loc_pair, X.contiguous(), pb.dst("A", part), W, 1)
                pb.barrier()
                X23c = pb.local("A").view(W * plan.m_loc, cw)                          # all rows of my columns
                ops._fused_pair_apply_push(cols, plan.sizes, ctx.col_pair, X23c, pb.dst("B", part), W, 3 if cw % 32 == 0 else 2)
                pb.barrier()
                ctx.save_for_backward(cols, X, X23c)
                return pb.local("B").view(W, plan.m_loc, cw)                            # my rows of every column block
            xb = comm.send_buffer(W * plan.m_loc * cw, X)            # symmetric (peer-mapped) send buffer or None
            X23s = ops._fused_pair_apply(cols, slab, ctx.loc_pair, X.contiguous(), chunk_out=W,
                                         out=None if xb is None else xb.view(W, plan.m_loc, cw))   # send layout
            X23c = comm.all_to_all(X23s.view(W, plan.m_loc, cw)).view(W * plan.m_loc, cw)     # all rows of my columns
            Yc = ops._fused_pair_apply(cols, plan.sizes, ctx.col_pair, X23c,
                                       out=None if xb is None else xb.view(W * plan.m_loc, cw))
            ctx.save_for_backward(cols, X, X23c)
            return comm.all_to_all(Yc.view(W, plan.m_loc, cw))                                # [W, m_loc, cw] blocks
        ctx.save_for_backward(cols, X)
        Y = _apply_local_axes(X, cols, plan)
        Yc = _to_cols(Y, plan, comm)
        Yc = ops.kron_axis_apply(Yc, cols[0], plan.sizes[0], 1, Yc.numel() // plan.sizes[0])
        return _to_rows(Yc, plan, comm).unsqueeze(0)

    @staticmethod
    def backward(ctx, gYb):
        plan, comm = ctx.plan, ctx.comm
        W = plan.world
        if ctx.fused:
            cols, X, X23c = ctx.saved_tensors
            d, gmax = cols.shape
            acc = torch.zeros(d, gmax, dtype=torch.float64, device=X.device)
            slab = [plan.g0_loc] + plan.sizes[1:]
            cw = X.shape[1] // W
            if ctx.pushed:
                pb = comm.push
                part = plan.m_loc * cw
                Cl = pb.local("C")
                if pb.c_pushed and gYb.data_ptr() == Cl.data_ptr():
                    pb.c_pushed = False                  # the Gram backward's GEMM already stored Z = L grad_Q on the owners
                else:
                    if pb.c_pushed:
                        raise RuntimeError("sharded Kronecker backward: the pushed gradient panel was replaced upstream")
                    Cl.view(W, plan.m_loc, cw).copy_(comm.all_to_all(gYb.contiguous()))
                Zc = Cl.view(W * plan.m_loc, cw)
                out = torch.zeros(2, 3, dtype=torch.float64, device=X.device)
                ops._fused_pair_grad_dir_push(cols, ctx.dirs, plan.sizes, ctx.col_pair, Zc, X23c, out[0], pb.dst("D", part), W)
                pb.barrier()
                ops._fused_pair_grad_dir(cols, ctx.dirs, slab, ctx.loc_pair, pb.local("D").view(W, plan.m_loc, cw), X, out[1],
                                         store=False, chunk_z=W)
                comm.allreduce_(out)
                gcols = ops._surrogate_col_grad(cols, ctx.dirs, ops._by_axis(out, [ctx.col_pair, ctx.loc_pair], plan.d), out[-1, 2])
                return gcols, None, None, None, None
            xb = comm.send_buffer(W * plan.m_loc * cw, X)
            if xb is not None and gYb.untyped_storage().data_ptr() != xb.untyped_storage().data_ptr():
                gYb = xb.view(W, plan.m_loc, cw).copy_(gYb)         # (the Gram backward normally writes it there itself)
            Zc = comm.all_to_all(gYb.contiguous()).view(W * plan.m_loc, cw)
            if ctx.dirs is not None:
                # directional form: 3 numbers per pair pass instead of two 32-entry column gradients; the partial sums
                # of the ranks (pair (0,1): my column block, pair (2,3): my row slab) meet in one 6-double all-reduce
                out = torch.zeros(2, 3, dtype=torch.float64, device=X.device)
                Z01c = ops._fused_pair_grad_dir(cols, ctx.dirs, plan.sizes, ctx.col_pair, Zc, X23c, out[0], store=True,
                                                zout=None if xb is None else xb.view(W * plan.m_loc, cw))
                if W > 1:
                    Z01b = comm.all_to_all(Z01c.view(W, plan.m_loc, cw))
                    ops._fused_pair_grad_dir(cols, ctx.dirs, slab, ctx.loc_pair, Z01b, X, out[1], store=False, chunk_z=W)
                    comm.allreduce_(out)
                else:
                    ops._fused_pair_grad_dir(cols, ctx.dirs, slab, ctx.loc_pair, Z01c, X, out[1], store=False)
                gcols = ops._surrogate_col_grad(cols, ctx.dirs, ops._by_axis(out, [ctx.col_pair, ctx.loc_pair], plan.d), out[-1, 2])
                return gcols, None, None, None, None
            Z01c = ops._fused_pair_grad(cols, plan.sizes, 0, Zc, X23c, acc, store=True,      # (full-gradient form: pairs (0,1), (2,3))
                                        zout=None if xb is None else xb.view(W * plan.m_loc, cw))   # axes 0, 1 (+ Z01)
            Z01b = comm.all_to_all(Z01c.view(W, plan.m_loc, cw))                               # chunked row layout
            ops._fused_pair_grad(cols, slab, 1, Z01b, X, acc, store=False, chunk_z=W)          # axes 2, 3
            comm.allreduce_(acc)
            return acc.to(cols.dtype), None, None, None, None
        gY = gYb[0]
        cols, X = ctx.saved_tensors
        d, gmax = cols.shape
        c = X.shape[1]
        acc = torch.zeros(d, gmax, dtype=torch.float64, device=X.device)
        axes = plan.local_axes(c)
        # suffix chain on X: S_i = T_{i+1} .. T_{d-1} X  (all slab-local)
        S = [None] * d
        S[d - 1] = X
        for (i, g, outer, inner) in reversed(axes):            # i = d-1 .. 1 -> S_{i-1} = T_i S_i
            S[i - 1] = ops.kron_axis_apply(S[i], cols[i], g, outer, inner)
        # axis 0 in the column-sharded layout: contraction, then the first step of the prefix chain on Z
        Zc = _to_cols(gY.contiguous(), plan, comm)
        S0c = _to_cols(S[0], plan, comm)
        inner0 = Zc.numel() // plan.sizes[0]
        ops.kron_axis_contract(Zc, S0c, plan.sizes[0], 1, inner0, acc[0])
        Pz = _to_rows(ops.kron_axis_apply(Zc, cols[0], plan.sizes[0], 1, inner0), plan, comm) if d > 1 else None
        for (i, g, outer, inner) in axes:
            ops.kron_axis_contract(Pz, S[i], g, outer, inner, acc[i])
            if i < d - 1:
                Pz = ops.kron_axis_apply(Pz, cols[i], g, outer, inner)
        comm.allreduce_(acc)
        return acc.to(cols.dtype), None, None, None, None


class _DualKronPushFn(torch.autograd.Function):
    """K L for the dual-layout model with the pushing kernels: L is also kept column-sharded (Lc [m, cw]: all rows of my
    columns), so both pair passes run on local data and only the LAST one's stores cross NVLink — into region B of every
    peer, which receives its rows of my column block.  Returns the column blocks [W, m_loc, cw] (region B).
    Backward: the Gram backward's panel GEMM has already stored Z = L grad_Q column-sharded in region C; both directional
    pair passes then run locally.  Two NVLink-bound kernels and two barriers per step instead of four and four."""

    @staticmethod
    def forward(ctx, cols, Lc, plan, comm, dirs):
        cols = cols.contiguous()
        ctx.plan, ctx.comm = plan, comm
        ctx.dirs = dirs.detach().to(cols.dtype).contiguous()
        W = plan.world
        cw = Lc.shape[1]
        pb = comm.push
        ctx.col_pair, ctx.loc_pair = ops.kron_pairs(plan.sizes, directional=True)
        X12 = ops._fused_pair_apply(cols, plan.sizes, ctx.loc_pair, Lc)
        ops._fused_pair_apply_push(cols, plan.sizes, ctx.col_pair, X12, pb.dst("B", plan.m_loc * cw), W, 3 if cw % 32 == 0 else 2)
        pb.barrier()
        ctx.save_for_backward(cols, Lc, X12)
        return pb.local("B").view(W, plan.m_loc, cw)

    @staticmethod
    def backward(ctx, gYb):
        plan, comm = ctx.plan, ctx.comm
        W = plan.world
        cols, Lc, X12 = ctx.saved_tensors
        cw = Lc.shape[1]
        pb = comm.push
        Cl = pb.local("C")
        if pb.c_pushed and gYb.data_ptr() == Cl.data_ptr():
            pb.c_pushed = False
        else:
            if pb.c_pushed:
                raise RuntimeError("sharded Kronecker backward: the pushed gradient panel was replaced upstream")
            Cl.view(W, plan.m_loc, cw).copy_(comm.all_to_all(gYb.contiguous()))
        Zc = Cl.view(W * plan.m_loc, cw)
        out = torch.zeros(2, 3, dtype=torch.float64, device=Lc.device)
        Z03 = ops._fused_pair_grad_dir(cols, ctx.dirs, plan.sizes, ctx.col_pair, Zc, X12, out[0], store=True)
        ops._fused_pair_grad_dir(cols, ctx.dirs, plan.sizes, ctx.loc_pair, Z03, Lc, out[1], store=False)
        comm.allreduce_(out)
        gcols = ops._surrogate_col_grad(cols, ctx.dirs, ops._by_axis(out, [ctx.col_pair, ctx.loc_pair], plan.d), out[-1, 2])
        return gcols, None, None, None, None


class _ShardedGramBlocksFn(torch.autograd.Function):
    """A_loc^T [B_0 | B_1 | ...] summed over ranks for column blocks Bb [nb, m_loc, cwb] (replicated r x (nb cwb)
    result); gradient w.r.t. the sharded blocks only, returned in the same block layout."""

    @staticmethod
    def forward(ctx, A, Bb, comm):
        ctx.save_for_backward(A)
        ctx.comm = comm
        ctx.nb, ctx.cwb = Bb.shape[0], Bb.shape[2]
        return comm.allreduce_(ops.gram_blocks(A, Bb, symmetric=True))      # L^T (K L): symmetric after the sum over ranks


    @staticmethod
    def backward(ctx, gG):
        (A,) = ctx.saved_tensors
        pb = ctx.comm.push_buffers(A.shape[0] * gG.shape[1], A) if (ctx.nb > 1 and ctx.nb == ctx.comm.world) else None
        if pb is not None and settings.kron_directional_grad.on():
            # Z = L grad_Q, column block j stored on rank j by the GEMM's epilogue (region C); what comes back is the
            # column-sharded panel the Kronecker backward needs, already in place
            ops.rmul_push(A, gG.contiguous(), pb.dst("C", A.shape[0] * ctx.cwb), ctx.nb, terms=ops._bwd_terms())
            pb.barrier()
            pb.c_pushed = True
            return None, pb.local("C").view(ctx.nb, A.shape[0], ctx.cwb), None
        xb = ctx.comm.send_buffer(A.shape[0] * gG.shape[1], A) if ctx.nb > 1 else None
        out = None if xb is None else xb.view(ctx.nb, A.shape[0], ctx.cwb)
        return None, ops.rmul_blocks(A, gG.contiguous(), ctx.nb, out=out, terms=ops._bwd_terms()), None


class _AllReduceGradFn(torch.autograd.Function):
    """Identity on a replicated tensor whose consumers are rank-local: the backward sums the per-rank gradients."""

    @staticmethod
    def forward(ctx, t, comm):
        ctx.comm = comm
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return ctx.comm.allreduce_(g.contiguous().clone()), None


class _ColsToRowBlocksFn(torch.autograd.Function):
    """Column-sharded panel [m, cw] (all rows of my columns) -> row-sharded column blocks [world, m_loc, cw] (my rows of
    every column block): one exchange; the backward is the inverse exchange."""

    @staticmethod
    def forward(ctx, Yc, plan, comm):
        ctx.plan, ctx.comm = plan, comm
        return comm.all_to_all(Yc.view(plan.world, plan.m_loc, Yc.shape[1]))

    @staticmethod
    def backward(ctx, gB):
        plan = ctx.plan
        return ctx.comm.all_to_all(gB.contiguous()).view(plan.world * plan.m_loc, gB.shape[2]), None, None


class _ShardedGramFn(torch.autograd.Function):
    """A_loc^T B_loc summed over ranks (replicated result); gradient w.r.t. the sharded B_loc only."""

    @staticmethod
    def forward(ctx, A, B, comm):
        ctx.save_for_backward(A)
        return comm.allreduce_(ops.gram(A, B))

    @staticmethod
    def backward(ctx, gG):
        (A,) = ctx.saved_tensors
        return None, ops._rmul(A, gG.contiguous(), ops._bwd_terms()), None


class _AllReduceFwdFn(torch.autograd.Function):
    """Sum of per-rank partials whose consumers are replicated: all-reduce forward, identity backward (every rank holds the
    full gradient of the sum already, and its partial enters the sum with weight one)."""

    @staticmethod
    def forward(ctx, t, comm):
        return comm.allreduce_(t.contiguous().clone())

    @staticmethod
    def backward(ctx, g):
        return g, None


class _ShardSliceFn(torch.autograd.Function):
    """Replicated full vector -> my slab; backward all-gathers the slab gradients so the replicated graph upstream
    receives the complete gradient on every rank."""

    @staticmethod
    def forward(ctx, full, plan, comm):
        ctx.plan, ctx.comm = plan, comm
        return full[plan.row0:plan.row0 + plan.m_loc].contiguous()

    @staticmethod
    def backward(ctx, g):
        return ctx.comm.allgather(g.contiguous()).reshape(-1, g.shape[-1]), None, None


def sharded_q_matvec(L_loc, KLb, v, comm):
    """w = (I + L^T K L) v for a replicated v [r, c] with row-sharded panels: every rank runs one pass over its slabs
    of K L (column blocks KLb [nb, m_loc, r / nb]) and L, then ONE all-reduce of r x c elements — the matmul closure
    of the CG solve (GPyTorch ``linear_cg``)."""
    nb, cwb = KLb.shape[0], KLb.shape[2]
    t = ops.panel_rmul(KLb[0], v[:cwb].contiguous())
    for j in range(1, nb):
        t = t + ops.panel_rmul(KLb[j], v[j * cwb:(j + 1) * cwb].contiguous())
    return v + comm.allreduce_(ops.gram(L_loc, t))


def sharded_cg_solve(L_loc, KLb, rhs, comm, tol=1e-2, max_iter=1000, check_every=4):
    """Solve (I + L^T K L) x = rhs by conjugate gradients with ``sharded_q_matvec``; GPyTorch ``linear_cg`` rules:
    columns L2-normalised first and rescaled at the end, stop when the mean residual norm < tol and at least
    min(10, max_iter - 1) iterations ran.  The r-sized recurrences are replicated (identical on every rank, so no
    further synchronisation); the residual is read back every ``check_every`` iterations.
    Returns (x, iterations, residual)."""
    with torch.no_grad():
        norm = rhs.norm(dim=0, keepdim=True).clamp_min(1e-10)
        b = rhs / norm
        x = torch.zeros_like(b)
        res = b.clone()
        p = res.clone()
        rz = (res * res).sum(0, keepdim=True)
        min_iter = min(10, max_iter - 1)
        it, resid = 0, 1.0
        while it < max_iter:
            Ap = sharded_q_matvec(L_loc, KLb, p, comm)
            alpha = rz / (p * Ap).sum(0, keepdim=True).clamp_min(1e-30)
            x = x + alpha * p
            res = res - alpha * Ap
            rz_new = (res * res).sum(0, keepdim=True)
            p = res + (rz_new / rz.clamp_min(1e-30)) * p
            rz = rz_new
            it += 1
            if it % check_every == 0 or it == max_iter:
                resid = float(res.norm(dim=0).mean())
                if it >= min_iter and resid < tol:
                    break
        return x * norm, it, resid


# ---------------------------------------------------------------------------------------------- the sharded model
class ShardedOnlineSKIRegression(torch.nn.Module):
    """Row-sharded counterpart of ``OnlineSKIRegression`` (Identity stem, one output, learnable noise)."""

    def __init__(self, init_x, init_y, lr, grid_size, grid_bound, comm=None, covar_module=None):
        super().__init__()
        self.comm = comm if comm is not None else Comm()
        d = init_x.shape[-1]
        assert init_y.ndim == 2 and init_y.shape[-1] == 1, "sharded path: one output"
        grid_bound = grid_bound + 1e-1                                           # online_ski_regression.py:26
        sizes = [grid_size] * d if isinstance(grid_size, int) else list(grid_size)
        base = covar_module if covar_module is not None else ScaleKernel(RBFKernel(ard_num_dims=d))
        self.covar_module = GridInterpolationKernel(base, grid_size=sizes, num_dims=d,
                                                    grid_bounds=torch.tensor([[-grid_bound, grid_bound]] * d)).to(init_x.device)
        self.likelihood = FNMGLikelihood(noise=torch.ones_like(init_y).t(), learn_additional_noise=True).to(init_x.device)
        self.plan = ShardPlan(sizes, self.comm.world, self.comm.rank)
        self.dtype = init_y.dtype
        self.gp_optimizer = torch.optim.Adam(self.parameters(), lr=lr)
        self._dual = settings.sharded_dual_layout.on() and self.comm.world > 1    # (one rank: the slab already holds all rows)
        self.Lc = None
        self._init_caches(init_x, init_y[:, 0], torch.ones_like(init_y[:, 0]))
        if self._dual:
            self.Lc = self._rows_to_cols(self.L_loc)           # [m, r / world]: all rows of my columns
        self._pieces = None
        if self.comm.world > 1 and init_x.is_cuda and _fused_ok(self.plan, self.L_loc):
            pushing = (settings.kron_directional_grad.on()
                       and self.comm.enable_push(self.L_loc.numel(), self.dtype, init_x.device) is not None)
            if not pushing:
                self.comm.enable_peer_exchange(self.L_loc.numel(), self.dtype, init_x.device)
        if self.comm.world > 1 and init_x.is_cuda and self.dtype == torch.float32:
            self.comm.enable_fast_allreduce(init_x.device)
        self._pending_root = None    # (p, C p^T) of a pre-started rank-q update (_prestart_root_update)
        self._graphs = None          # opt-in CUDA-graph replay: enable_cuda_graphs()
        self._n_t = None             # device-side copy of num_data (graph mode)

    # ---- state
    def _stencils(self, x):
        idx, val = self.covar_module._compute_grid(x)
        return self.plan.localize(idx, val.detach())

    def _init_caches(self, X, y, D):
        plan, comm = self.plan, self.comm
        idx, val = self._stencils(X)
        n0 = X.shape[0]
        self.response_cache = (y * y / D).sum()
        self.D_logdet = D.log().sum()
        self.num_data = n0
        self.b_loc = ops.left_t_interp(idx, val, (y / D).unsqueeze(-1), plan.m_loc)
        # the interpolation cache b = W^T D^-1 y is an m-vector (4 MB at m = 2^20): every rank keeps it whole as well (each
        # update scatters q s entries), so building K b needs no all-gather
        gidx, gval = self.covar_module._compute_grid(X)
        self.b_full = ops.left_t_interp(gidx, gval.detach(), (y / D).unsqueeze(-1), plan.m)
        vval = val / D.clamp_min(1e-7).sqrt().unsqueeze(-1)
        max_rank = settings.max_root_decomposition_size.value()
        n1 = min(n0, max_rank)
        eye = torch.eye(n1, dtype=self.dtype, device=X.device)
        V1 = ops.left_t_interp(idx[:n1], vval[:n1], eye, plan.m_loc)
        lam, U = torch.linalg.eigh(comm.allreduce_(ops.gram(V1, V1)))
        tol = 1e-10 if self.dtype == torch.float64 else 1e-5
        keep = lam > tol * lam.max()
        lam, U = lam[keep].flip(0), U[:, keep].flip(1)
        r_eff = lam.numel()
        # multiple of 32 x ranks: the column-sharded layout of the fused kernels needs whole 16-column tiles per rank, the
        # chunked tensor-core panel GEMM whole 32-column groups (1 rank: 16, as the single-GPU model)
        mult = 16 if comm.world == 1 else 32 * comm.world
        r = ((r_eff + mult - 1) // mult) * mult
        Upad = torch.zeros(n1, r, dtype=self.dtype, device=X.device)
        Upad[:, :r_eff] = U
        scale = torch.zeros(r, dtype=self.dtype, device=X.device)
        scale[:r_eff] = 1.0 / lam
        self.L_loc = ops.panel_rmul(V1, Upad)
        self.B_loc = (self.L_loc * scale).contiguous()
        if n0 > n1:
            self._fold_in(idx[n1:], vval[n1:])

    def _fold_in(self, idx_l, vval_l, chunk=4096):
        """Batched projected update for the initial points beyond the root rank — the sharded form of
        ``UpdatedRootLazyTensor.fold_in_sparse``: P = B^T V (all-reduced partial row gathers), M = I + P P^T factored
        once (replicated), one local panel GEMM per panel."""
        r = self.B_loc.shape[1]
        M = torch.eye(r, dtype=torch.float64, device=self.B_loc.device)
        for s0 in range(0, idx_l.shape[0], chunk):
            Pc = self.comm.allreduce_(ops.left_interp(idx_l[s0:s0 + chunk], vval_l[s0:s0 + chunk].contiguous(),
                                                      self.B_loc)).double()
            M = M + Pc.t() @ Pc
        F = torch.linalg.cholesky(M)
        Finv_t = torch.linalg.solve_triangular(F, torch.eye(r, dtype=torch.float64, device=M.device), upper=False).t()
        self.L_loc = ops.panel_rmul(self.L_loc, F.to(self.dtype).contiguous())
        self.B_loc = ops.panel_rmul(self.B_loc, Finv_t.to(self.dtype).contiguous())

    def _rows_to_cols(self, P_loc):
        """Row slab [m_loc, r] -> column block [m, r / world] (one all-to-all; used once, at construction)."""
        W, plan = self.comm.world, self.plan
        cw = P_loc.shape[1] // W
        send = P_loc.view(plan.m_loc, W, cw).permute(1, 0, 2).contiguous()
        return self.comm.all_to_all(send).reshape(W * plan.m_loc, cw).contiguous()

    def _root_update(self, idx_l, vval_l):
        """collect_vector (updated_root_lazy_tensor.py:69-119), symmetric-square-root form, panels updated in place."""
        pend, self._pending_root = self._pending_root, None
        for s0 in range(0, idx_l.shape[0], 32):
            if pend is not None:
                p, CpT = pend                      # inverse-root half already under way on the side stream (_prestart_root_update)
                ops.join_side(self.L_loc.device)
            else:
                pT = self.comm.allreduce_(ops.left_interp(idx_l[s0:s0 + 32], vval_l[s0:s0 + 32], self.B_loc))
                p = pT.t().contiguous()
                C, Cp = _sym_factors(p)
                CpT = C @ p.t()
            if self.Lc is not None:
                # the same update on the column-sharded copy: L p for ALL rows — a by-product of the rank-q launch on the
                # row slab, all-gathered as an m x q vector — then Lc += (L p) (C p^T)[:, my columns]
                cw = self.Lc.shape[1]
                if pend is not None:
                    _, t_loc = ops.panel_lowrank_update1_(self.L_loc, p, CpT, return_t=True)
                else:
                    _, _, t_loc = ops.panel_lowrank_update2_(self.L_loc, self.B_loc, p, CpT, Cp @ p.t(), return_t=True)
                t_full = self.comm.allgather(t_loc).reshape(self.plan.m, p.shape[1])
                ops.panel_outer_add_(self.Lc, t_full, CpT[:, self.comm.rank * cw:(self.comm.rank + 1) * cw].contiguous())
            elif pend is not None:
                ops.panel_lowrank_update1_(self.L_loc, p, CpT)
            else:
                ops.panel_lowrank_update2_(self.L_loc, self.B_loc, p, CpT, Cp @ p.t())

    def _prestart_root_update(self, x):
        """settings.overlap_root_update: projection and factors of the rank-q update that ends this step now, the in-place
        update of the inverse-root slab on a side stream (background launch) under the hyper-parameter step, which never
        reads it.  ``_root_update`` then joins and updates the root slab.  One block of q <= 32 points only."""
        if not (settings.overlap_root_update.on() and ops.overlap_capable(x) and 1 <= x.shape[0] <= 32) \
                or self._pending_root is not None:
            return
        with torch.no_grad():
            idx_l, val_l = self._stencils(x)                 # D = 1 in the streaming update (online_ski_regression.py:122)
            pT = self.comm.allreduce_(ops.left_interp(idx_l, val_l, self.B_loc))
            p = pT.t().contiguous()
            C, Cp = _sym_factors(p)
            CpT, CppT = (C @ p.t()).contiguous(), (Cp @ p.t()).contiguous()
            with ops.side_section(self.B_loc.device):
                ops.panel_lowrank_update1_(self.B_loc, p, CppT)
            self._pending_root = (p, CpT)

    # ---- pieces with grad (Kuu / sigma^2, K L, Q, K b, c)  — batched_fixed_noise_online_gp.py:334-366
    def _noise(self):
        return self.likelihood.second_noise_covar.noise.to(self.dtype).reshape(())

    def _build_pieces(self):
        plan, comm = self.plan, self.comm
        cols = self.covar_module.base_kernel.grid_columns(self.covar_module.grid).to(self.dtype)
        noise = self._noise()
        scale = torch.cat([(1.0 / noise).reshape(1, 1), torch.ones(plan.d - 1, 1, dtype=self.dtype, device=cols.device)])
        cols = cols * scale                                                       # Kuu / sigma^2 (:340)
        dirs = self.covar_module.base_kernel.grid_column_dirs(self.covar_module.grid) \
            if settings.kron_directional_grad.on() else None
        if self.Lc is not None and comm.push_buffers(self.L_loc.numel(), self.L_loc) is not None and dirs is not None:
            KL = _DualKronPushFn.apply(cols, self.Lc, plan, comm, dirs)
        elif self.Lc is not None:
            # dual layout: K L on the column-sharded copy is the ordinary (single-device) Kronecker MVM with its own
            # autograd; one exchange brings it to row-sharded column blocks, the backward sends the gradient back
            cw = self.Lc.shape[1]
            xb = comm.send_buffer(plan.m * cw, self.Lc)
            # (the surrogate column gradient of the directional backward is linear in the per-rank partial sums, so
            #  summing it over the ranks in _AllReduceGradFn gives the surrogate of the complete sums)
            KLc = ops.kron_toeplitz_matmul(_AllReduceGradFn.apply(cols, comm), plan.sizes, self.Lc, dirs=dirs,
                                           out=None if xb is None else xb.view(plan.m, cw))
            KL = _ColsToRowBlocksFn.apply(KLc, plan, comm)
        else:
            KL = _ShardedKronFn.apply(cols, self.L_loc, plan, comm, dirs)         # :348  column blocks [nb, m_loc, r / nb]
        r = self.L_loc.shape[1]
        b_full = self.b_full
        Kb_full = ops.kron_toeplitz_matmul(cols, plan.sizes, b_full)              # :366
        Kb = _ShardSliceFn.apply(Kb_full, plan, comm)
        overlap = settings.overlap_root_update.on() and ops.overlap_capable(Kb)
        if overlap:
            # c = L^T K b: the HBM-bound pass over the row slab (and, in the backward, L g_c) runs as a background launch on a
            # side stream, under the tensor-bound Gram (backward: panel GEMM) issued next on this one; the sum over ranks
            # stays on this stream, after the Gram's own
            with ops.side_section(Kb.device):
                c_part = ops.gram(self.L_loc, Kb)
        Q = _ShardedGramBlocksFn.apply(self.L_loc, KL, comm) + torch.eye(r, dtype=self.dtype, device=KL.device)   # :352-355
        if overlap:
            ops.join_side(Kb.device)
            c = _AllReduceFwdFn.apply(c_part, comm)
        else:
            c = _ShardedGramFn.apply(self.L_loc, Kb, comm)                        # :360-361
        Lq, _ = torch.linalg.cholesky_ex(Q, check_errors=False)      # Q >= I: no host-side info check (no sync)
        self._pieces = dict(cols=cols, KL=KL, Q=Q, Lq=Lq, Kb_full=Kb_full, Kb=Kb, c=c, b_full=b_full, noise=noise)
        return self._pieces

    def pieces(self):
        return self._pieces if self._pieces is not None else self._build_pieces()

    # ---- evaluate / predict  (online_ski_regression.py:56-78)
    def predict(self, x):
        P = self.pieces()
        plan, comm = self.plan, self.comm
        with torch.no_grad():
            idx, val = self.covar_module._compute_grid(x)
            val = val.detach()
            idx_l, val_l = plan.localize(idx, val)
            KLb = P["KL"].detach()
            nb, cwb = KLb.shape[0], KLb.shape[2]
            a = self._q_solve(P, P["c"].detach())
            # W* (K b - K L Q^-1 c) (:206-210, :376) without materialising the m-vector: the stencil rows of K b and of K L
            # are gathered once (the K L rows also serve the variance) and meet in ONE all-reduce of q x (1 + r) numbers
            G = comm.allreduce_(torch.cat([ops.left_interp(idx_l, val_l, P["Kb"].detach())]
                                          + [ops.left_interp(idx_l, val_l, KLb[j]) for j in range(nb)], dim=1))
            T = G[:, 1:].t()
            mean = G[:, :1] - T.t() @ a
            q = x.shape[0]
            Wt = ops.left_t_interp(idx, val, torch.eye(q, dtype=self.dtype, device=x.device), plan.m)
            c1 = ops.left_interp(idx, val, ops.kron_toeplitz_matmul(P["cols"].detach(), plan.sizes, Wt))
            cov = (c1 - T.t() @ self._q_solve(P, T)) * P["noise"]                  # :222-228
            var = cov.diagonal().unsqueeze(-1) + P["noise"]                        # predict(): + second_noise
        return mean, var

    def _q_solve(self, P, rhs):
        """Q^-1 rhs: Cholesky up to ``max_cholesky_size`` (every shipped config), otherwise the sharded CG driver — one
        pass over the local panel slabs and one all-reduce per iteration (eval-only, like the single-GPU CG path)."""
        if rhs.shape[0] <= settings.max_cholesky_size.value():
            return torch.cholesky_solve(rhs, P["Lq"])
        tol = settings.eval_cg_tolerance.value()
        x, iters, resid = sharded_cg_solve(self.L_loc, P["KL"].detach(), rhs, self.comm, tol=tol,
                                           max_iter=settings.max_cg_iterations.value())
        if resid > tol:
            warnings.warn(f"CG terminated in {iters} iterations with average residual norm {resid:.3e} which is larger "
                          f"than the tolerance of {tol} specified by eval_cg_tolerance.", RuntimeWarning)
        self.last_cg = (iters, resid)
        return x

    def _evaluate_stats(self, x, y):
        mean, var = self.predict(x)
        rmse = (mean - y).pow(2).mean().sqrt()
        nll = -torch.distributions.Normal(mean, var.sqrt(), validate_args=False).log_prob(y).mean()
        return torch.stack([rmse, nll])

    def evaluate(self, x, y):
        if self._graph_usable(x):
            return self._evaluate_graphed(x, y)
        self._graph_phase(None)
        with settings.defer_interp_bounds_check(x.is_cuda):
            rmse, nll = self._evaluate_stats(x, y).tolist()          # one device->host read
        ops.flush_bounds_checks()
        return rmse, nll

    # ---- Woodbury MLL (batched_woodbury_marginal_log_likelihood.py:19-52)
    def mll(self):
        P = self.pieces()
        Lq = P["Lq"]
        half = torch.linalg.solve_triangular(Lq, P["c"], upper=False)
        inner_qform = (half * half).sum()
        logdet = 2.0 * Lq.diagonal().log().sum()
        if settings.skip_logdet_forward.on():
            logdet = logdet - logdet.detach()
        inducing_qform = (P["b_full"] * P["Kb_full"]).sum()
        inv_quad = (self.response_cache - inducing_qform + inner_qform) / P["noise"]
        n = self.num_data if self._n_t is None else self._n_t
        final = n * math.log(2 * math.pi) + n * P["noise"].log()
        return -0.5 * (inv_quad + logdet + self.D_logdet + final) / n

    # ---- update (online_ski_regression.py:113-146)
    def _hyper_step(self):
        self.gp_optimizer.zero_grad()
        with settings.skip_logdet_forward(True):
            loss = -self.mll()
        loss.backward()
        self.gp_optimizer.step()
        self._pieces = None
        return loss.detach()

    def update(self, x, y):
        if self._graph_usable(x) and self._graphs.phase == "evaluated":
            return self._update_graphed(x, y)
        self._graph_phase(None)
        self._prestart_root_update(x)
        loss = self._hyper_step()
        with torch.no_grad():
            self.condition_on_observations(x, y[:, 0], torch.ones_like(y[:, 0]))
        return 0.0, loss.item()          # read back with the conditioning kernels already queued

    def condition_on_observations(self, x, y, D):
        idx_l, val_l = self._stencils(x)
        self.response_cache.add_((y * y / D).sum())           # in place: graph replays must see the running values
        self.D_logdet.add_(D.log().sum())
        ops.scatter_add_(self.b_loc, idx_l, val_l, (y / D).unsqueeze(-1))
        gidx, gval = self.covar_module._compute_grid(x)
        ops.scatter_add_(self.b_full, gidx, gval.detach(), (y / D).unsqueeze(-1))
        self._root_update(idx_l, val_l / D.clamp_min(1e-7).sqrt().unsqueeze(-1))
        self.num_data += x.shape[0]
        if self._n_t is not None:
            self._n_t.add_(x.shape[0])
        self._pieces = None

    # ---- CUDA-graph replay (online_gp_b200/graphs.py); every rank takes the same path by construction
    def enable_cuda_graphs(self, enabled=True, warmup_calls=2):
        self._graphs = StepGraphs(warmup_calls) if enabled else None
        return self

    def _graph_phase(self, phase):
        if self._graphs is not None:
            self._graphs.phase = phase

    def _graph_usable(self, x):
        G = self._graphs
        return not (G is None or G.failed or not x.is_cuda or (G.q is not None and x.shape[0] != G.q))

    def _evaluate_graphed(self, x, y):
        G = self._graphs
        if G.eval is None and G.warm > 0:
            G.warm -= 1
            self._graphs = None
            try:
                return self.evaluate(x, y)
            finally:
                self._graphs = G
        if G.q is None:
            G.setup(x, y)
            self._n_t = torch.full((), float(self.num_data), dtype=self.dtype, device=x.device)
            make_adam_capturable(self.gp_optimizer)
        G.load(x, y)
        if G.eval is None:
            self._pieces = None
            try:
                G.eval = G.capture(lambda: self._evaluate_stats(G.x, G.y))
            except Exception as err:            # noqa: BLE001
                G.fail(err)
                self._pieces = None
                return self.evaluate(x, y)
        rmse, nll = G.replay(G.eval)
        G.phase = "evaluated"
        return rmse, nll

    def _update_graphed(self, x, y):
        G = self._graphs
        n_before = self.num_data
        G.load(x, y)
        if G.upd is None:
            def body():
                self._prestart_root_update(G.x)
                loss = self._hyper_step()
                with torch.no_grad():
                    self.condition_on_observations(G.x, G.y[:, 0], torch.ones_like(G.y[:, 0]))
                return loss
            try:
                G.upd = G.capture(body)
            except Exception as err:            # noqa: BLE001
                self.num_data = n_before
                G.fail(err)
                G.phase = None
                self._pieces = None
                return self.update(x, y)
        try:
            (loss,) = G.replay(G.upd)
        finally:                     # keep the host bookkeeping in step with the device even if the replay raises
            G.phase = None
            self.num_data = n_before + x.shape[0]
            self._pieces = None
        return 0.0, loss

    @property
    def graph_launches(self):
        return 0 if self._graphs is None else self._graphs.launches
