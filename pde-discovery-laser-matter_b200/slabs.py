"""Time-slab sharding of the fused path: one process per GPU, contiguous frame
ranges per rank, a one-frame trailing halo (the forward u_t of ks2d:1511 needs u(t+1)), and a
single small all-reduce of the per-fold statistics.  ``torch.distributed`` is the plumbing
(NCCL over NVLink on GPUs; gloo in the CPU tests); the data path has exactly these two
exchanges.  Slab edges are aligned to ``block_t`` so no block straddles two ranks.

Also here: ``fit_streamed`` for host-resident stacks (slabs copied from pinned host memory on
a copy stream while the previous slab is reduced on the compute stream).
"""

from __future__ import annotations

import numpy as np

from . import _lib as L


def slab_bounds(n_row_frames: int, block_t: int, world: int, allow_empty: bool = False):
    """Row-frame ranges [lo, hi) per rank: whole t-blocks, as even as possible; a ragged last
    t-block (ks2d:384) stays with the last rank.  More ranks than t-blocks is an error (an empty slab has no first
    frame to hand to its left neighbour) unless ``allow_empty`` -- for callers that drop the empty ranges themselves."""
    n_tb = -(-n_row_frames // block_t)
    if world > n_tb and not allow_empty:
        raise ValueError(f"{world} ranks for {n_tb} t-block(s) of {block_t} frame(s): use at most {n_tb} ranks")
    base, extra = divmod(n_tb, world)
    out, tb = [], 0
    for r in range(world):
        k = base + (1 if r < extra else 0)
        lo, hi = min(n_row_frames, tb * block_t), min(n_row_frames, (tb + k) * block_t)
        out.append((lo, max(lo, hi)))
        tb += k
    return out


class PeerComm:
    """The C-ABI communicator (pg_comm_init / pg_comm_barrier / pg_allreduce_stats / pg_halo_exchange,
    include/pdegram.h) over NVLink peer memory.

    torch symmetric memory (CUDA VMM allocations mapped into every rank of the group) is only the ALLOCATOR and the
    exchange of the mappings; what runs are the library's own kernels and copy-engine transfers on plain pointers:

    * ``allreduce(stats)``: ONE launch; every rank stores its vector into every peer, waits for all flags and sums in
      rank order, so the result is the same bits on every rank and from run to run (NCCL promises neither).
    * ``slab(shape)``: a time slab allocated IN symmetric memory: the next rank's first frame can then be pulled
      straight out of its slab, with no publishing copy.
    * ``pull_halo(U_local)``: barrier (one warp, before K1), then on a side stream a copy engine pulls the next rank's
      first frame into ``U_local[-1]`` and a stream memory operation raises a flag; returns ``(flag_ptr, epoch)`` for
      ``ops.fd_lib_gram(..., halo=...)``: K1 is ONE launch that starts immediately and polls the flag only before it
      loads the trailing frame (which the last t-block alone reads).

    Raises when symmetric memory cannot be set up (CPU, one GPU, no peer access); callers keep ``PeerHalo``'s send/recv
    path for that."""

    def __init__(self, group=None, device=None):
        import ctypes as C

        import torch.distributed as dist

        torch = L.torch_cuda()
        import torch.distributed._symmetric_memory as symm_mem

        self.lib = L.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > L.PG_COMM_MAX_RANKS:
            raise L.PdeGramError(f"at most {L.PG_COMM_MAX_RANKS} ranks")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._symm = symm_mem
        nbytes = int(self.lib.pg_comm_workspace_bytes())
        self.ws = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.ws.zero_()
        torch.cuda.synchronize()
        self.ws_hdl = symm_mem.rendezvous(self.ws, self.group)
        ptrs = self._peer_ptrs(self.ws_hdl, self.ws)
        arr = (C.c_void_p * self.world)(*ptrs)
        h = C.c_void_p()
        L.check(self.lib.pg_comm_init(self.rank, self.world, arr, C.byref(h)))
        self.h = h
        dist.barrier(group)                 # every workspace is zeroed before anybody signals into it
        torch.cuda.synchronize()
        self.copy_stream = torch.cuda.Stream()
        self._slabs = {}                    # data_ptr of a symmetric slab -> pointer of the NEXT rank's copy
        self._pub = None                    # publish buffers for slabs that are not symmetric

    def _peer_ptrs(self, hdl, t):
        try:
            return [int(x) for x in hdl.buffer_ptrs]
        except Exception:
            torch = L.torch_cuda()
            return [int(hdl.get_buffer(q, tuple(t.shape), t.dtype).data_ptr()) for q in range(self.world)]

    def slab(self, shape, dtype=None):
        """A (T, A0, A1) stack in symmetric memory (collective: every rank calls it with the same shape)."""
        import torch.distributed as dist

        torch = L.torch_cuda()
        shape = tuple(int(x) for x in shape)
        numel = int(np.prod(shape))
        # slabs of a ragged partition differ by a t-block: every rank allocates the largest one (symmetric allocations
        # are matched by size) and works on a view of its own shape
        n = torch.tensor([numel], dtype=torch.int64, device=self.device)
        dist.all_reduce(n, op=dist.ReduceOp.MAX, group=self.group)
        base = self._symm.empty(int(n.item()), dtype=dtype or torch.float64, device=self.device)
        hdl = self._symm.rendezvous(base, self.group)
        ptrs = self._peer_ptrs(hdl, base)
        t = base[:numel].view(shape)
        self._slabs[int(t.data_ptr())] = (hdl, ptrs[self.rank + 1] if self.rank < self.world - 1 else None, base)
        return t

    def release(self, t):
        self._slabs.pop(int(t.data_ptr()), None)

    def barrier(self):
        L.check(self.lib.pg_comm_barrier(self.h, L.stream_ptr()))

    def allreduce(self, stats):
        if not stats.is_contiguous():
            raise L.PdeGramError("allreduce needs a contiguous statistics tensor")
        L.check(self.lib.pg_allreduce_stats(self.h, L.ptr(stats), stats.numel(), L.stream_ptr()))
        return stats

    def errors(self) -> int:
        return int(self.lib.pg_comm_errors(self.h))

    def pull_halo(self, U_local):
        """Start pulling the next rank's first frame into U_local[-1]; returns (flag_ptr, epoch) or None on the last
        rank.  Stream-ordered: everything queued on the current stream before this call (e.g. the producer of
        U_local) happens before the peers read it."""
        import ctypes as C

        torch = L.torch_cuda()
        frame = U_local[0]
        nbytes = frame.numel() * frame.element_size()
        entry = self._slabs.get(int(U_local.data_ptr()))
        if entry is None:
            # not a symmetric slab: publish the first frame in a symmetric buffer (two alternate, so a slower
            # neighbour may still be pulling the previous one)
            if self._pub is None or self._pub[0].shape[1:] != frame.shape:
                buf = self._symm.empty((2,) + tuple(frame.shape), dtype=frame.dtype, device=self.device)
                hdl = self._symm.rendezvous(buf, self.group)
                self._pub = (buf, hdl, self._peer_ptrs(hdl, buf), 0)
            buf, hdl, ptrs, k = self._pub
            self._pub = (buf, hdl, ptrs, k + 1)
            b = k & 1
            buf[b].copy_(frame)
            src = ptrs[self.rank + 1] + b * nbytes if self.rank < self.world - 1 else None
        else:
            src = entry[1]
        self.barrier()
        if src is None:
            return None
        ev = torch.cuda.Event()
        ev.record()
        self.copy_stream.wait_event(ev)
        flag, epoch = C.c_void_p(), C.c_uint32()
        L.check(self.lib.pg_halo_exchange(self.h, L.ptr(U_local[-1]), src, nbytes, int(self.copy_stream.cuda_stream),
                                          C.byref(flag), C.byref(epoch)))
        return int(flag.value), int(epoch.value)

    def close(self):
        if getattr(self, "h", None):
            self.lib.pg_comm_destroy(self.h)
            self.h = None


class PeerHalo:
    """Trailing-halo exchange through NVLink peer memory, with no SM involvement.

    K1 is a persistent kernel that fills every SM (one CTA, ~218 KB of shared memory each), so a NCCL
    send/recv kernel launched beside it cannot be scheduled until K1 retires: the exchange would serialise
    behind the bulk launch instead of hiding under it.  Here every rank publishes its first frame in a
    symmetric-memory buffer (``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped into every
    peer over NVSwitch), one small barrier kernel orders the publication BEFORE K1 starts, and the frame of
    rank r+1 is then pulled by a copy engine (cudaMemcpyAsync on a side stream) while K1 owns the SMs.
    Two buffers alternate, so a rank never overwrites a frame a slower neighbour may still be pulling.

    Falls back to NCCL/gloo send-recv (``exchange_halo_begin``) when symmetric memory cannot be set up
    (CPU tests, single GPU, unsupported platform); ``mode`` says which path is active.
    """

    def __init__(self, frame_shape, group=None, device=None, peer_memory=True):
        import torch.distributed as dist

        self.group, self.mode, self.k = group, "send_recv", 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.done = None
        if self.world == 1:
            self.mode = "none"
            return
        if not peer_memory:          # A/B comparisons: NCCL / gloo point-to-point
            self.why = "disabled by the caller"
            return
        try:
            torch = L.torch_cuda()
            import torch.distributed._symmetric_memory as symm_mem

            shape = (2,) + tuple(int(x) for x in frame_shape)
            self.buf = symm_mem.empty(shape, dtype=torch.float64, device=device or torch.device("cuda", torch.cuda.current_device()))
            self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
            self.peer = self.hdl.get_buffer(self.rank + 1, shape, torch.float64) if self.rank < self.world - 1 else None
            self.stream = torch.cuda.Stream()
            if shape[1] % 8 == 0 and shape[2] % 8 == 0:
                # compressed exchange (begin_block_means): (8, 8) block means of the first frame, 1/64 of its bytes
                mshape = (2, shape[1] // 8, shape[2] // 8)
                self.bm = symm_mem.empty(mshape, dtype=torch.float64, device=self.buf.device)
                self.bm_hdl = symm_mem.rendezvous(self.bm, group if group is not None else dist.group.WORLD)
                self.bm_peer = (self.bm_hdl.get_buffer(self.rank + 1, mshape, torch.float64)
                                if self.rank < self.world - 1 else None)
                self.tail = torch.empty(mshape, dtype=torch.float64, device=self.buf.device)
            self.mode = "peer_memory"
        except Exception as exc:  # symmetric memory unavailable: keep the send/recv path
            self.why = repr(exc)

    def begin(self, U_local):
        """Start filling U_local[-1] with the next rank's first frame; returns a token for end()."""
        if self.mode == "none":
            return None
        if self.mode == "send_recv":
            return exchange_halo_begin(U_local, self.group)
        torch = L.torch_cuda()
        b = self.k & 1
        self.k += 1
        self.buf[b].copy_(U_local[0])          # publish my first frame (local copy)
        self.hdl.barrier(channel=0)            # every rank has published; tiny kernel, stream-ordered before K1
        if self.peer is None:
            return None
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            U_local[-1].copy_(self.peer[b], non_blocking=True)   # copy engine over NVLink
            done = torch.cuda.Event()
            done.record(self.stream)
        return done

    def begin_block_means(self, U_local, means_fn=None):
        """Compressed exchange for (bt, 8m, 8n) blocks of the KS dialect: the forward time difference of a block
        telescopes to a difference of block sums, so a rank only needs the (8, 8) block MEANS of its neighbour's first
        frame (2 MB instead of 134 MB at 4096^2) and K1 runs as ONE launch over the whole slab
        (``ops.fd_lib_gram(..., trailing_block_means=...)``) instead of bulk + a tail launch that waits for the frame.
        Returns (token for end(), means or None on the last rank).  ``means_fn`` stands in for the CUDA kernel in
        the CPU tests."""
        if self.mode == "none":
            return None, None
        if means_fn is None:
            from . import ops

            means_fn = ops.frame_block_means
        if self.mode == "send_recv":
            import torch.distributed as dist

            mine = means_fn(U_local[0])
            out = mine.new_empty(mine.shape) if self.rank < self.world - 1 else None
            ops_ = []
            if self.rank > 0:
                ops_.append(dist.P2POp(dist.isend, mine, self.rank - 1, self.group))
            if out is not None:
                ops_.append(dist.P2POp(dist.irecv, out, self.rank + 1, self.group))
            reqs = dist.batch_isend_irecv(ops_) if ops_ else []
            self._keep = mine                      # the send buffer must outlive the request
            return reqs, out
        torch = L.torch_cuda()
        b = self.k & 1
        self.k += 1
        means_fn(U_local[0], out=self.bm[b])       # publish: the kernel writes into the symmetric buffer
        self.bm_hdl.barrier(channel=0)
        if self.bm_peer is None:
            return None, None
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self.tail[b].copy_(self.bm_peer[b], non_blocking=True)   # copy engine over NVLink
            done = torch.cuda.Event()
            done.record(self.stream)
        return done, self.tail[b]

    def end(self, token):
        """Make the current stream wait for the halo frame."""
        if token is None:
            return
        if self.mode == "send_recv":
            exchange_halo_end(token)
        else:
            L.torch_cuda().cuda.current_stream().wait_event(token)


def exchange_halo_begin(U_local, group=None):
    """Start filling the trailing halo frame U_local[-1] with frame 0 of the next rank's slab.

    Rank r sends its first frame to rank r-1 and receives rank r+1's first frame; the last
    rank keeps its own trailing frame (it is part of the global stack).  Returns the pending
    requests; the transfer runs on the communication stream and overlaps whatever the caller
    launches next, as long as that does not read U_local[-1]."""
    import torch.distributed as dist

    if not dist.is_initialized():
        return []
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return []
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, U_local[0], rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.irecv, U_local[-1], rank + 1, group))
    return dist.batch_isend_irecv(ops) if ops else []


def exchange_halo_end(reqs):
    """Make the current stream wait for the halo transfer started by exchange_halo_begin."""
    for req in reqs:
        req.wait()


def exchange_halo(U_local, group=None):
    """Blocking form: begin + end."""
    exchange_halo_end(exchange_halo_begin(U_local, group))
    return U_local


def allreduce_stats(stats, group=None):
    """Sum the per-fold statistics over ranks, in place (one all-reduce of n_folds*S doubles)."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def block_means_halo_ok(shape, dialect, block) -> bool:
    """Whether the trailing halo can travel as (8, 8) block means (pg_fd_lib_gram_tail's layout rule)."""
    _, A0, A1 = (int(x) for x in shape)
    return (dialect == L.FD_KS_PERIODIC and block[1] % 8 == 0 and block[2] % 8 == 0 and A0 % 8 == 0 and A1 % 8 == 0
            and A1 >= 128)


def halo_flag_ok(shape, dialect, block) -> bool:
    """Whether K1 can run as one launch that polls the halo flag (pg_fd_lib_gram_halo's layout rule)."""
    _, A0, A1 = (int(x) for x in shape)
    return (dialect == L.FD_KS_PERIODIC and int(block[1]) == 8 and int(block[2]) == 8 and A0 % 8 == 0 and A1 % 8 == 0
            and A1 >= 128)


def check_world(n_row_frames: int, block_t: int, world: int):
    """Every rank must own at least one t-block: an empty slab would publish a first frame it does not have (its
    U_local[0] is its own, not yet received, halo frame) and its left neighbour would difference against garbage."""
    n_tb = -(-int(n_row_frames) // int(block_t))
    if world > n_tb:
        raise ValueError(f"{world} ranks for {n_tb} t-block(s): shrink the group to at most {n_tb} ranks "
                         "(every rank needs at least one t-block)")


def sharded_stats(U_local, d0, d1, dt, *, dialect, library, block=(1, 1, 1), fold_of_frame=None, fold_of_row=None,
                  n_folds=1, variant=L.VARIANT_AUTO, group=None, halo=True, stats_fn=None, peer_halo=None,
                  means_fn=None, block_means_halo=None, comm=None):
    """Per-rank K1 over this rank's slab (own frames + trailing halo frame) followed by the
    all-reduce.  ``stats_fn`` lets the CPU tests stand in for the CUDA kernel.

    ``comm`` (a ``PeerComm``): the NVLink peer-memory path -- barrier, copy-engine pull of the halo frame behind a
    flag, ONE K1 launch that polls the flag before its last t-block, one-launch rank-ordered all-reduce.

    The halo frame is only read by the LAST t-block of the slab, so the exchange is started first,
    K1 runs on everything before that t-block while the frame is in flight, and only the small tail
    launch waits for it (the transfer is hidden; statistics are additive over time slabs)."""
    if U_local.shape[0] < 2:
        raise ValueError("a slab needs at least one row frame and its trailing halo frame (world > number of t-blocks?)")
    if comm is not None and stats_fn is None:
        from . import ops

        kw = dict(dialect=dialect, library=library, block=block, n_folds=n_folds, variant=variant,
                  fold_of_frame=fold_of_frame, fold_of_row=fold_of_row)
        tok = comm.pull_halo(U_local) if halo else None
        if tok is not None and not (halo_flag_ok(U_local.shape, dialect, block) and variant != L.VARIANT_GENERIC):
            L.torch_cuda().cuda.current_stream().wait_stream(comm.copy_stream)   # layout without the polling kernel
            tok = None
        return comm.allreduce(ops.fd_lib_gram(U_local, d0, d1, dt, halo=tok, **kw))
    if block_means_halo is None:
        # Measured on B200 (bench.py, PG_HALO): with NVLink peer memory the whole frame is pulled by a copy engine under
        # K1 and costs nothing, so the compressed form only saves the short tail launch and pays for a publish kernel
        # before K1 (a wash); with NCCL / gloo send-recv the frame exchange serialises behind the persistent K1
        # (+0.6 ms per step at 4096^2), so the 64x smaller message wins.
        block_means_halo = peer_halo is not None and peer_halo.mode == "send_recv"
    if block_means_halo and peer_halo is not None and halo and peer_halo.mode != "none" \
            and (stats_fn is None or means_fn is not None) and block_means_halo_ok(U_local.shape, dialect, block) \
            and variant != L.VARIANT_GENERIC:
        # compressed halo: block means of the neighbour's first frame, ONE K1 launch over the whole slab
        token, tail = peer_halo.begin_block_means(U_local, means_fn)
        peer_halo.end(token)
        if stats_fn is not None:
            return allreduce_stats(stats_fn(U_local, tail), group)
        from . import ops

        stats = ops.fd_lib_gram(U_local, d0, d1, dt, dialect=dialect, library=library, block=block, n_folds=n_folds,
                                variant=variant, fold_of_frame=fold_of_frame, fold_of_row=fold_of_row,
                                trailing_block_means=tail)
        return allreduce_stats(stats, group)
    if peer_halo is not None and halo:
        # NVLink peer-memory exchange (PeerHalo): same begin / end protocol
        token = peer_halo.begin(U_local)
        reqs = [token] if token is not None else []
        finish = lambda r: peer_halo.end(r[0]) if r else None   # noqa: E731
    else:
        reqs = exchange_halo_begin(U_local, group) if halo else []
        finish = exchange_halo_end
    if stats_fn is not None:
        finish(reqs)
        return allreduce_stats(stats_fn(U_local), group)
    from . import ops

    bt = int(block[0])
    rows = U_local.shape[0] - 1
    cut = ((rows - 1) // bt) * bt          # first frame of the last (possibly ragged) t-block
    kw = dict(dialect=dialect, library=library, block=block, n_folds=n_folds, variant=variant)
    if not reqs or cut <= 0 or fold_of_row is not None:
        finish(reqs)
        stats = ops.fd_lib_gram(U_local, d0, d1, dt, fold_of_frame=fold_of_frame, fold_of_row=fold_of_row, **kw)
    else:
        fof = None if fold_of_frame is None else ops._dev(fold_of_frame)
        stats = ops.fd_lib_gram(U_local[:cut + 1], d0, d1, dt, fold_of_frame=None if fof is None else fof[:cut], **kw)
        finish(reqs)
        ops.stats_accumulate(stats, ops.fd_lib_gram(U_local[cut:], d0, d1, dt, fold_of_frame=None if fof is None else fof[cut:], **kw))
    return allreduce_stats(stats, group)


def fit_streamed(U_host, d0, d1, dt, *, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8),
                 fold_of_frame=None, n_folds=1, slab_frames=96, variant=L.VARIANT_AUTO, buffers=None):
    """Statistics of a HOST-resident stack: slabs of ``slab_frames`` row frames (+1 halo frame)
    are copied host->device on a copy stream into two device buffers while the compute stream
    runs K1 on the previous slab; per-slab statistics are summed on the device.

    U_host: (T, A0, A1) float64 torch tensor in pinned memory (DMA straight from it), or a NumPy array / pageable
    tensor, whose slabs go through the pinned staging buffers of ``_xfer`` (multi-threaded host copy + DMA).
    Returns the [n_folds][S] statistics tensor (device).
    """
    torch = L.torch_cuda()
    from . import _xfer, ops

    staged = None
    if isinstance(U_host, np.ndarray):
        staged = np.ascontiguousarray(U_host, dtype=np.float64)
        U_host = torch.from_numpy(staged)
    elif not U_host.is_pinned():
        staged = U_host.contiguous().numpy()
    T, A0, A1 = U_host.shape
    bt = int(block[0])
    slab_frames = max(bt, (slab_frames // bt) * bt)
    p = L.LIB_WIDTH[library]
    total = torch.zeros((n_folds, L.stats_len(p)), dtype=torch.float64, device="cuda")
    if buffers is None:
        buffers = [torch.empty((slab_frames + 1, A0, A1), dtype=torch.float64, device="cuda") for _ in range(2)]
    fof = None if fold_of_frame is None else torch.as_tensor(np.asarray(fold_of_frame), dtype=torch.int32).cuda()
    compute = torch.cuda.current_stream()
    copy = torch.cuda.Stream()
    filled = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    starts = list(range(0, T - 1, slab_frames))
    for k, lo in enumerate(starts):
        hi = min(T - 1, lo + slab_frames)
        b = k % 2
        with torch.cuda.stream(copy):
            if k >= 2:
                copy.wait_event(freed[b])
            if staged is None:
                buffers[b][: hi - lo + 1].copy_(U_host[lo:hi + 1], non_blocking=True)
            else:
                _xfer.to_device(staged[lo:hi + 1], out=buffers[b][: hi - lo + 1])
            filled[b].record(copy)
        compute.wait_event(filled[b])
        s = ops.fd_lib_gram(buffers[b][: hi - lo + 1], d0, d1, dt, dialect=dialect, library=library, block=block,
                            fold_of_frame=None if fof is None else fof[lo:hi], n_folds=n_folds, variant=variant)
        ops.stats_accumulate(total, s)
        freed[b].record(compute)
    return total
