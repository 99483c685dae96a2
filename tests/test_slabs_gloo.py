"""Host logic of the multi-GPU path on CPU: world_size-2 gloo processes exercise the slab
partition, the one-frame halo exchange and the statistics all-reduce of pde_b200.slabs.  The
per-rank kernel is stood in for by the oracle (test-only), so what is pinned here is that
"sum over ranks of slab statistics == statistics of the whole stack"."""

import os
import socket

import numpy as np
import pytest

from helpers import assert_stats_close, ks_rows
from oracle import gram


def test_slab_bounds():
    from pde_b200.slabs import slab_bounds

    assert slab_bounds(1023, 3, 1) == [(0, 1023)]
    b = slab_bounds(1023, 3, 8)
    assert b[0][0] == 0 and b[-1][1] == 1023 and all(lo % 3 == 0 for lo, _ in b)
    assert all(b[k][1] == b[k + 1][0] for k in range(7))
    assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 3
    assert slab_bounds(10, 3, 2) == [(0, 6), (6, 10)]            # ragged last t-block stays last
    assert slab_bounds(4, 3, 4, allow_empty=True) == [(0, 3), (3, 4), (4, 4), (4, 4)]  # more ranks than t-blocks
    with pytest.raises(ValueError):      # ... is an error by default: an empty slab has no first frame to publish
        slab_bounds(4, 3, 4)


def _worker_too_many_ranks(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from pde_b200 import _lib as L
    from pde_b200 import slabs

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 3 row frames in blocks of 3 = ONE t-block for two ranks: both the partition and the per-rank entry refuse
        out = []
        try:
            slabs.slab_bounds(3, 3, world)
        except ValueError as exc:
            out.append("bounds:" + str(exc)[:20])
        try:
            slabs.check_world(3, 3, world)
        except ValueError as exc:
            out.append("check")
        U_local = torch.zeros((1, 8, 128), dtype=torch.float64)       # an empty slab: only its halo frame
        try:
            slabs.sharded_stats(U_local, 0.5, 0.5, 1e-2, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8),
                                stats_fn=lambda u: None)
        except ValueError:
            out.append("sharded")
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_more_ranks_than_t_blocks_is_rejected():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_too_many_ranks, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert len(res[r]) == 3 and res[r][1:] == ["check", "sharded"], res


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker_means(rank, world, port, q):
    """Compressed halo (block means of the neighbour's first frame) over gloo: the trailing frame stays a placeholder."""
    import torch
    import torch.distributed as dist

    from pde_b200 import _lib as L
    from pde_b200 import slabs

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        block = (3, 8, 8)
        U = np.random.default_rng(1).standard_normal((13, 16, 128))
        dx, dy, DT = 0.5, 0.25, 1e-2
        lo, hi = slabs.slab_bounds(U.shape[0] - 1, block[0], world)[rank]
        U_local = torch.from_numpy(U[lo:hi + 1].copy())
        if rank < world - 1:
            U_local[-1].zero_()

        def means_fn(frame, out=None):
            m = frame.reshape(frame.shape[0] // 8, 8, frame.shape[1] // 8, 8).mean(dim=(1, 3))
            return m if out is None else out.copy_(m)

        def stats_fn(Ul, tail=None):
            a = Ul.numpy().copy()
            if tail is not None:      # what the kernel does: only the block sums of the trailing frame enter (u_t telescopes)
                a[-1] = np.kron(tail.numpy(), np.ones((8, 8)))
            names, X, y = ks_rows(a, dx, dy, DT, "true", False, block)
            return torch.from_numpy(gram.pack_stats(X, y))[None].clone()

        peer = slabs.PeerHalo(U.shape[1:])
        assert peer.mode == "send_recv"
        s = slabs.sharded_stats(U_local, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=block,
                                stats_fn=stats_fn, means_fn=means_fn, peer_halo=peer)
        if rank < world - 1:
            assert float(U_local[-1].abs().max()) == 0.0      # the frame itself never travelled
        q.put((rank, s.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_block_means_halo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_means, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    U = np.random.default_rng(1).standard_normal((13, 16, 128))
    names, X, y = ks_rows(U, 0.5, 0.25, 1e-2, "true", False, (3, 8, 8))
    assert np.array_equal(res[0], res[1])
    assert_stats_close(res[0][0], gram.pack_stats(X, y), 3, rtol=1e-12)


def _worker(rank, world, port, block, q, use_peer=False):
    import torch
    import torch.distributed as dist

    from pde_b200 import _lib as L
    from pde_b200 import slabs

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        U = rng.standard_normal((14, 16, 24))
        dx, dy, DT = 0.5, 0.25, 1e-2
        lo, hi = slabs.slab_bounds(U.shape[0] - 1, block[0], world)[rank]
        U_local = torch.from_numpy(U[lo:hi + 1].copy())
        if rank < world - 1:
            U_local[-1].zero_()  # the halo frame must come from the neighbour

        def stats_fn(Ul):
            names, X, y = ks_rows(Ul.numpy(), dx, dy, DT, "rich", False, block)
            return torch.from_numpy(gram.pack_stats(X, y))[None].clone()

        peer = None
        if use_peer:
            # no CUDA / symmetric memory here: PeerHalo must say so and take the send/recv path
            peer = slabs.PeerHalo(U.shape[1:])
            assert peer.mode == "send_recv" and peer.world == world
        s = slabs.sharded_stats(U_local, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_RICH, block=block,
                                stats_fn=stats_fn, peer_halo=peer)
        assert torch.equal(U_local[-1], torch.from_numpy(U[hi]))
        q.put((rank, s.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("block,use_peer", [((1, 1, 1), False), ((3, 8, 8), False), ((3, 8, 8), True)])
def test_two_rank_halo_and_allreduce(block, use_peer):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, block, q, use_peer)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    U = np.random.default_rng(0).standard_normal((14, 16, 24))
    names, X, y = ks_rows(U, 0.5, 0.25, 1e-2, "rich", False, block)
    ref = gram.pack_stats(X, y)
    assert np.array_equal(res[0], res[1])        # every rank holds the reduced statistics
    assert_stats_close(res[0][0], ref, 9, rtol=1e-12)


def test_peer_halo_single_process_is_a_no_op():
    from pde_b200.slabs import PeerHalo

    h = PeerHalo((4, 4))
    assert h.mode == "none" and h.begin(None) is None
    h.end(None)
