"""Hardware parity of the SHARDED path on two real GPUs: NCCL + NVLink peer memory.

Spawns two ranks when the box has at least two GPUs (skipped otherwise; the host logic of the same functions is
covered on CPU by tests/test_slabs_gloo.py).  Every rank generates the whole small stack locally (the generator is
deterministic in the global frame index), so both the frame a halo mechanism must deliver and the unsharded
statistics are at hand: for each mechanism the pulled frame must be bit-identical and the all-reduced statistics
must equal the unsharded ones to 1e-12 (the oracle comparison of the unsharded statistics is the rest of the suite).
Mechanisms: PeerComm (flag-polled single K1 launch + one-launch peer all-reduce) with the slab in symmetric memory
and with a published frame, PeerHalo frame pull, PeerHalo block means, NCCL send/recv of the frame and of the means.
"""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), NCCL_DEBUG="WARN")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from helpers import stats_scales
        from pde_b200 import _lib as L
        from pde_b200 import ops, slabs

        out = {}
        D0, D1, DT = 0.5, 0.4, 1e-3
        for case, (A0, A1, tb_rank, ragged, lib, p) in {
            "true_64x256": (64, 256, 4, 2, L.LIB_KS_TRUE, 3),
            "rich_72x384_ragged_rows": (72, 384, 3, 1, L.LIB_KS_RICH, 9),
        }.items():
            g_rows = world * tb_rank * 3 + ragged
            whole = ops.synth_field(g_rows + 1, A0, A1, t_offset=0, T_total=64, seed=5, noise=0.05)
            split = int(0.7 * g_rows) // 3 * 3
            fof = (np.arange(g_rows) >= split).astype(np.int32)
            kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=(3, 8, 8), n_folds=2)
            ref = ops.fd_lib_gram(whole, D0, D1, DT, fold_of_frame=fof, **kw).cpu().numpy()
            lo, hi = slabs.slab_bounds(g_rows, 3, world)[rank]
            comm = slabs.PeerComm()
            for mode in ("flag", "flag_pub", "peer", "means", "send_recv", "send_recv_means"):
                Ul = comm.slab((hi - lo + 1, A0, A1)) if mode == "flag" else \
                    torch.empty((hi - lo + 1, A0, A1), dtype=torch.float64, device="cuda")
                Ul.copy_(whole[lo:hi + 1])
                if rank < world - 1:
                    Ul[-1].fill_(float("nan"))           # a frame that never arrives must not pass
                for rep in range(3):                     # repeated calls: epochs, alternating buffers
                    if rank < world - 1:
                        Ul[-1].fill_(float("nan"))
                    if mode.startswith("flag"):
                        got = slabs.sharded_stats(Ul, D0, D1, DT, fold_of_frame=fof[lo:hi], comm=comm, **kw)
                    else:
                        ph = slabs.PeerHalo((A0, A1), peer_memory=not mode.startswith("send_recv")) if rep == 0 else ph
                        got = slabs.sharded_stats(Ul, D0, D1, DT, fold_of_frame=fof[lo:hi], peer_halo=ph,
                                                  block_means_halo=mode.endswith("means"), **kw)
                    torch.cuda.synchronize()
                    g = got.cpu().numpy()
                    err = max(float((np.abs(g[f] - ref[f]) / np.maximum(stats_scales(ref[f], p), 1e-300)).max()) for f in range(2))
                    frame_ok = mode.endswith("means") or rank == world - 1 or bool(torch.equal(Ul[-1], whole[hi]))
                    out[(case, mode, rep)] = (err, frame_ok, bool(np.array_equal(g[:, 0], ref[:, 0])))
                if mode == "flag":
                    comm.release(Ul)
            # the one-launch all-reduce: same bits on every rank, equal to the rank-ordered sum
            v = torch.arange(40, dtype=torch.float64, device="cuda") * (rank + 1) * 0.1 + 1e-3 * rank
            mine = v.clone()
            comm.allreduce(v)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            want = parts[0].clone()
            for k in range(1, world):
                want += parts[k]
            out[(case, "allreduce_bitwise")] = bool(torch.equal(v, want))
            out[(case, "comm_errors")] = comm.errors()
            comm.close()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_parity_every_mechanism():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank in range(2):
        for key, val in res[rank].items():
            if key[1] == "allreduce_bitwise":
                assert val is True, (rank, key)
            elif key[1] == "comm_errors":
                assert val == 0, (rank, key)
            else:
                err, frame_ok, n_ok = val
                assert frame_ok, f"rank {rank} {key}: the halo frame that arrived is not the true frame"
                assert n_ok and err <= 1e-12, f"rank {rank} {key}: sharded vs unsharded statistics differ by {err:.3e}"
