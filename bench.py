#!/usr/bin/env python
"""Benchmark of the fused FD + library + Gram hot path (BASELINE.json metric: grid points/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over the rank's synthetic stack: (N>1: barrier + one-frame halo
pull over NVLink) -> K1 pg_fd_lib_gram over the whole slab with two time-holdout folds -> (N>1: one
all-reduce of the 2 x S statistics) -> K3 batched STRidge (the reference's 5 x 6 sweep with held-out
r2/rmse).  Workload of the contract line (config.workload): BASELINE configs[3], a 2048 x 2048 x 1024
float64 stack per GPU in the KS-2D dialect with the reference's default true dictionary (p = 3) and
(3, 8, 8) block averaging (configs[1]'s estimator at configs[3]'s size).  N > 1 shards contiguous time
slabs, one per GPU, weak scaling (the global stack is N x 1023 row frames + 1).

The same line carries, measured in the same run:
  parity    (N > 1) hardware parity of the sharded path: pulled halo frame bit-identical to the true
            frame, all-reduced statistics == rank-ordered sum of the per-slab statistics, and sharded ==
            unsharded == oracle on a small stack through every halo mechanism
  c5        BASELINE configs[4]: ONE 4096 x 4096 x 2048 stack cut into N time slabs (strong scaling),
            with the single-GPU streamed baseline timed by rank 0 in the same run -> speedup
  roofline, cpu_baseline, e2e, gpu_launches, clocks, variants, patch_ensemble, reference_configs

`--impl reference` times the reference's own CPU path (its unmodified functions from baseline/_ref
when staged, else the NumPy port in oracle/) on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "grid points/sec (fused FD+library+Gram)"
UNIT = "grid-points/s"
D0 = D1 = 0.5
DT = 1e-3
BLOCK = (3, 8, 8)
C5_SIZE, C5_FRAMES = 4096, 2048      # BASELINE configs[4]

def fp64_per_point():
    """fp64 instructions per grid point of the K1 specialisations, counted per opcode (DADD/DFMA/DMUL) on the ncu source
    pages and committed as profiles/fp64_per_point.json (tools/summarize_profiles.py).  The fp64 roofline of a variant
    is points/s x this / (SMs x 64 per clock x SM clock)."""
    f = ROOT / "profiles" / "fp64_per_point.json"
    if not f.exists():
        return {}
    return {k: v["fp64_per_point"] for k, v in json.loads(f.read_text())["kernels"].items()}


def workload_text(args, world):
    return (f"c4: synthetic {args.size}x{args.size}x{args.frames} float64 stack per GPU, KS periodic dialect, "
            f"true dictionary p=3, block average (3,8,8), 2 time-holdout folds (70/30), 5x6 STRidge sweep; "
            f"{world} time slab(s)")


# ----------------------------------------------------------------------------- CPU arms (reference arm / cpu_baseline)
_CPU = {}  # sample stack shared with the forked worker processes (no pickling)


def _cpu_slab_stats(a):
    lo, hi, dictionary = a
    from oracle import gram, ks2d as O

    sl = _CPU["U"][lo:hi + 1]
    if dictionary == "true":
        names, terms = O.build_dictionary_true(sl[:-1], D0, D1)
    else:
        names, terms = O.build_dictionary(sl[:-1], D0, D1)
    X, y = O.build_blockwise_dataset((sl[1:] - sl[:-1]) / DT, terms, names, block_t=BLOCK[0], block_x=BLOCK[1],
                                     block_y=BLOCK[2])
    return gram.pack_stats(X, y)


def cpu_port_step(U, workers, pool, dictionary="true"):
    """The reference algorithm (oracle port) for one pass over sample U: per-time-slab FD +
    dictionary + block means + Gram in `workers` processes, then the STRidge sweep."""
    from oracle import gram, ks2d as O
    from pde_b200.slabs import slab_bounds

    bounds = [b for b in slab_bounds(U.shape[0] - 1, BLOCK[0], workers, allow_empty=True) if b[1] > b[0]]
    parts = pool.map(_cpu_slab_stats, [(lo, hi, dictionary) for lo, hi in bounds])
    n_tr = max(1, int(0.7 * len(parts)))
    s_tr, s_te = sum(parts[:n_tr]), sum(parts[n_tr:]) if len(parts) > n_tr else sum(parts)
    p = 3 if dictionary == "true" else 9
    return gram.ks_fit_from_stats(s_tr, s_te, p, alphas=O.GRID_ALPHAS, thresholds=O.GRID_THRESHOLDS,
                                  const_cols=() if dictionary == "true" else (0,))


def cpu_sample(frames=25, size=512, seed=0):
    """Bounded sample of the workload for the CPU arm: same generator family, same dialect /
    library / block, (frames x size x size) float64."""
    rng = np.random.default_rng(seed)
    i = np.arange(size) * (2 * np.pi / size)
    t = np.arange(frames) * (2 * np.pi / 1024)
    a, b, s = i[None, :, None], i[None, None, :], t[:, None, None]
    U = (0.5 * np.sin(3 * a + 2 * b - 5 * s) + 0.3 * np.sin(7 * a - 4 * b + 3 * s + 0.7)
         + 0.15 * np.sin(13 * a + 11 * b - 9 * s + 1.9) + 0.05 * np.cos(29 * a - 17 * b + 2 * s))
    return U + 0.05 * (rng.random(U.shape) - 0.5)


def time_cpu_port(steps, warmup, frames, size):
    import multiprocessing as mp

    workers = max(1, min(os.cpu_count() or 1, (frames - 1) // BLOCK[0]))
    U = cpu_sample(frames, size)
    _CPU["U"] = U
    with mp.get_context("fork").Pool(workers) as pool:
        for _ in range(warmup):
            cpu_port_step(U, workers, pool)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_port_step(U, workers, pool)
        dt = (time.perf_counter() - t0) / steps
    pts = frames * size * size
    return pts / dt, dt, workers, f"{frames}x{size}x{size} float64 sub-stack, same dialect/library/block, {workers} worker processes over time slabs"


def reference_step(ks, U):
    """One pass of the hot path through the UNMODIFIED reference functions (baseline/_ref): forward u_t,
    build_dictionary_true (ks2d:1063), build_blockwise_dataset (ks2d:358), time-holdout split, train-RMS scale
    (ks2d:1647-1655) and the 5 x 6 stridge sweep with the reference's arg-max key (ks2d:1720-1743)."""
    Uf, Ut = U[:-1], (U[1:] - U[:-1]) / DT
    names, terms = ks.build_dictionary_true(Uf, dx=D0, dy=D1)
    X, y = ks.build_blockwise_dataset(Ut, terms, names, block_t=BLOCK[0], block_x=BLOCK[1], block_y=BLOCK[2])
    n_tb = -(-(U.shape[0] - 1) // BLOCK[0])
    rows_tb = len(y) // n_tb
    split = rows_tb * max(1, int(0.7 * n_tb))
    Xtr, ytr, Xte, yte = X[:split], y[:split], X[split:], y[split:]
    scale = np.sqrt(np.mean(Xtr ** 2, axis=0)) + 1e-12
    best = None
    for a in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2):
        for t in (1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5):
            c = ks.stridge(Xtr / scale, ytr, alpha=a, threshold=t, max_iter=25) / scale
            pred = Xte @ c
            key = (ks.r2_score(yte, pred), -int(np.sum(np.abs(c) > 0)), -ks.rmse(yte, pred))
            if best is None or key > best[0]:
                best = (key, a, t, c)
    return dict(alpha=best[1], threshold=best[2], coeffs=best[3])


def time_reference(steps, warmup, frames, size):
    """The reference's own functions, single process (the reference has no parallelism; NumPy's BLAS threads are
    whatever the box gives it)."""
    from oracle import refload

    ks = refload.load("ks2d")
    U = cpu_sample(frames, size)
    for _ in range(warmup):
        reference_step(ks, U)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = reference_step(ks, U)
    dt = (time.perf_counter() - t0) / steps
    return frames * size * size / dt, dt, out, (f"{frames}x{size}x{size} float64 sub-stack through the UNMODIFIED reference functions "
                                                 "(build_dictionary_true, build_blockwise_dataset, stridge; baseline/_ref), 1 process")


def cpu_baseline_object(steps=2, warmup=0):
    """cpu_baseline of the contract: the reference itself when baseline/_ref is staged (kind "reference"), the
    multi-process NumPy port always (kind "port"; reported beside it)."""
    from oracle import refload

    v, _, workers, sample = time_cpu_port(3, 1, 97, 1024)
    port = {"value": v, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample}
    if not refload.available("ks2d"):
        return port
    vr, dtr, _, sample_r = time_reference(steps, warmup, 49, 512)
    return {"value": vr, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample_r, "s_per_step": dtr,
            "port": port}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refload

    steps, warmup = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    t_budget = time.perf_counter()
    pv, pdt, workers, psample = time_cpu_port(min(max(1, args.steps), 5), max(0, min(args.warmup, 2)), *args.port_sample)
    port = {"value": pv, "unit": UNIT, "cores": workers, "kind": "port", "sample": psample, "ms_per_step": pdt * 1e3}
    if refload.available("ks2d") and not args.port_only:
        value, dt, _, sample = time_reference(steps, warmup, *args.ref_sample)
        cpu = {"value": value, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample}
        note = ("the reference is CPU-only, single-process Python: its UNMODIFIED functions (baseline/_ref) are timed on a "
                "bounded sample; `port` is the vectorised NumPy restatement (oracle/) with one process per host core")
    else:
        value, dt, cpu, steps, warmup = pv, pdt, dict(port), min(max(1, args.steps), 5), max(0, min(args.warmup, 2))
        cpu.pop("ms_per_step")
        note = "reference scripts not staged (baseline/_ref absent): timed as the NumPy port (oracle/) on a bounded sample"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args, args.gpus), "note": note},
        "cpu_baseline": cpu, "port": port,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_budget,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """SM / memory clock, power, temperature and throttle reasons DURING the timed region, sampled through NVML
    (nvidia_ml_py, ~2 ms period; `nvidia-smi -lms` cannot resolve a sub-second region)."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.rows, self.run, self.thread, self.h = [], False, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.n = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # pragma: no cover
            self.err = repr(exc)

    def start(self):
        if self.h is None:
            return
        self.run = True
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        n = self.n
        while self.run:
            try:
                row = [time.perf_counter(), n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM),
                       n.nvmlDeviceGetPowerUsage(self.h) / 1e3, n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)]
                try:
                    row.append(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_MEM))
                    row.append(n.nvmlDeviceGetTemperature(self.h, n.NVML_TEMPERATURE_GPU))
                except Exception:
                    row += [None, None]
                self.rows.append(row)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.run = False
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["sampler disabled"]}
        self.thread.join(timeout=2)
        if os.environ.get("PG_BENCH_CLOCK_TRACE"):
            t0 = self.rows[0][0] if self.rows else 0.0
            print("clock trace (ms, sm MHz, W, reasons, mem MHz, degC):",
                  [(round((r[0] - t0) * 1e3, 1), r[1], round(r[2]), hex(r[3]), r[4], r[5]) for r in self.rows], file=sys.stderr)
        sm = [r[1] for r in self.rows]
        mem = [r[4] for r in self.rows if r[4] is not None]
        temp = [r[5] for r in self.rows if r[5] is not None]
        bits = 0
        for r in self.rows:
            bits |= r[3]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": self.max_sm, "reasons": sorted(v for k, v in self.REASONS.items() if bits & k),
                "power_w_max": max((r[2] for r in self.rows), default=None), "samples": len(sm),
                "mem_mhz_min": min(mem) if mem else None, "mem_mhz_max": max(mem) if mem else None,
                "temp_c_max": max(temp) if temp else None}


class _StdoutToStderr:
    """NCCL prints its version banner to fd 1 from C; the contract is ONE JSON line on stdout.  Route fd 1
    to fd 2 while the communicators are being created."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def stats_rel_err(got, ref, p):
    """(max error against each entry's Cauchy-Schwarz scale, max ENTRYWISE relative error |got-ref|/|ref| over the
    entries with |ref| above 1e-6 of their scale).  The first is the norm the 1e-10 gate uses (a cross term such as
    sum u_x*u_y cancels to ~0 and has no own magnitude); the second is the north star's wording taken literally,
    reported next to it."""
    got, ref = np.asarray(got, dtype=np.float64).ravel(), np.asarray(ref, dtype=np.float64).ravel()
    S = 3 + 2 * p + p * (p + 1) // 2
    out_cs, out_ew = 0.0, 0.0
    for k in range(len(ref) // S):
        g, r = got[k * S:(k + 1) * S], ref[k * S:(k + 1) * S]
        n, syy = r[0], r[2]
        G = np.zeros((p, p))
        G[np.triu_indices(p)] = r[3 + 2 * p:]
        d = np.sqrt(np.maximum(np.diag(G), 0))
        scale = np.concatenate([[max(n, 1.0), np.sqrt(n * syy), syy], np.sqrt(n) * d, np.sqrt(syy) * d, np.outer(d, d)[np.triu_indices(p)]])
        scale = np.maximum(scale, 1e-300)
        err = np.abs(g - r)
        out_cs = max(out_cs, float((err / scale).max()))
        sig = np.abs(r) > 1e-6 * scale
        if sig.any():
            out_ew = max(out_ew, float((err[sig] / np.abs(r[sig])).max()))
    return out_cs, out_ew


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import pde_b200
    from pde_b200 import _lib as L
    from pde_b200 import ks2d as K
    from pde_b200 import ops, slabs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # PG_HALO selects the exchange mechanism of the sharded path (A/B comparisons; profiles/README.md):
    #   flag (default)  PeerComm: slab in symmetric memory, copy engine pulls the neighbour's first frame behind a flag,
    #                   ONE K1 launch that polls it, one-launch rank-ordered all-reduce through peer memory
    #   peer            round 1: frame published + pulled through peer memory, bulk + tail K1 launches, NCCL all-reduce
    #   means           (8, 8) block means of the frame through peer memory, one K1 launch, NCCL all-reduce
    #   send_recv / send_recv_means   the same two through NCCL point-to-point
    halo_kind = os.environ.get("PG_HALO", "flag")
    comm, halo, comm_why = None, None, None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        with _StdoutToStderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)          # creates the communicator (and prints NCCL's banner) now
            torch.cuda.synchronize()
            if halo_kind == "flag":
                try:
                    comm = slabs.PeerComm()
                except Exception as exc:   # no symmetric memory on this platform: round-1 path
                    comm_why = repr(exc)[:200]
                ok = torch.tensor([1 if comm is not None else 0], device="cuda")
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if not int(ok.item()):
                    comm, halo_kind = None, "peer"
    lib = pde_b200.load()
    alphas = ops._dev(np.array(K.GRID_ALPHAS))
    thrs = ops._dev(np.array(K.GRID_THRESHOLDS))
    names = K.TRUE_NAMES

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    class Job:
        """One workload on this rank: `g_rows` global row frames (+1 trailing frame) of an A x A field cut into
        `w` time slabs; this rank (`r`) owns slab r and walks it in `passes` sub-slabs through ONE buffer (a slab
        that does not fit the GPU: c5 on one GPU; the on-device generator refills the buffer between passes and is
        excluded from the timing)."""

        def __init__(self, A, g_rows, w, r, comm_=None, halo_=None, kind="none"):
            self.A, self.g_rows, self.w, self.r, self.comm, self.halo, self.kind = A, g_rows, w, r, comm_, halo_, kind
            lo, hi = slabs.slab_bounds(g_rows, BLOCK[0], w)[r]
            self.lo, self.hi = lo, hi
            limit = float(os.environ.get("PG_BENCH_PASS_BYTES", 150e9))  # per-GPU buffer budget (env: test hook)
            passes = 1
            while True:
                sub = [(lo + a, lo + b) for a, b in slabs.slab_bounds(hi - lo, BLOCK[0], passes)]
                if (max(b - a for a, b in sub) + 1) * A * A * 8 <= limit:
                    break
                passes += 1
            if passes > 1 and w > 1:
                raise RuntimeError("a slab that needs several passes AND a halo exchange is not supported")
            self.sub, self.passes = sub, passes
            self.Tbuf = max(b - a for a, b in sub) + 1
            self.U = comm_.slab((self.Tbuf, A, A)) if (comm_ is not None and kind == "flag") else \
                torch.empty((self.Tbuf, A, A), dtype=torch.float64, device="cuda")
            split = int(0.7 * g_rows) // BLOCK[0] * BLOCK[0]
            self.fof = [torch.from_numpy((np.arange(a, b) >= split).astype(np.int32)).cuda() for a, b in sub]
            self.k1_ev, self.pass_ev, self.tail_ev = [], [], []
            self.fill(0)

        def fill(self, ps, true_halo=False):
            """(Re)generate sub-slab `ps`; the trailing frame of the rank's LAST sub-slab comes from the next rank by
            halo exchange (zeroed here), every other one from the generator."""
            a, b = self.sub[ps]
            last = ps == self.passes - 1 and self.r < self.w - 1 and not true_halo
            n = b - a + 1
            ops.synth_field(n - 1 if last else n, self.A, self.A, t_offset=a, T_total=1024, seed=0, noise=0.05, out=self.U[:n])
            if last:
                self.U[n - 1].zero_()

        def view(self, ps):
            a, b = self.sub[ps]
            return self.U[:b - a + 1]

        def k1(self, ps, library, block, variant, record):
            """K1 over sub-slab ps (with the halo exchange on the last one); returns the [2][S] statistics."""
            V, fof = self.view(ps), self.fof[ps]
            kw = dict(dialect=L.FD_KS_PERIODIC, library=library, block=block, n_folds=2, variant=variant)
            exch = self.w > 1 and ps == self.passes - 1
            want_ev = record or self.passes > 1
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if want_ev else None
            if exch and self.kind == "flag":
                if ev:
                    ev[2].record()
                tok = self.comm.pull_halo(V)                     # barrier (one warp) + copy-engine pull behind a flag
                if ev:
                    ev[0].record()
                s = ops.fd_lib_gram(V, D0, D1, DT, fold_of_frame=fof, halo=tok, **kw)   # ONE launch, polls the flag
                if ev:
                    ev[1].record()
                    if record:
                        self.k1_ev.append([(ev[0], ev[1])])
                    self.pass_ev.append((ev[2], ev[1]))
                return s
            if exch and self.kind.endswith("means"):
                if ev:
                    ev[2].record()
                token, tail = self.halo.begin_block_means(V)
                self.halo.end(token)
                if ev:
                    ev[0].record()
                s = ops.fd_lib_gram(V, D0, D1, DT, fold_of_frame=fof, trailing_block_means=tail, **kw)
                if ev:
                    ev[1].record()
                    if record:
                        self.k1_ev.append([(ev[0], ev[1])])
                    self.pass_ev.append((ev[2], ev[1]))
                return s
            # whole frame through PeerHalo (peer memory or send/recv): bulk launch while the frame is in flight, then the
            # last t-block (the only reader of the halo frame) as a short tail launch
            token = self.halo.begin(V) if exch else None
            rows = V.shape[0] - 1
            cut = ((rows - 1) // block[0]) * block[0] if token is not None else rows
            if ev:
                ev[0].record()
            s = ops.fd_lib_gram(V[:cut + 1], D0, D1, DT, fold_of_frame=fof[:cut], **kw)
            if ev:
                ev[1].record()
            if token is not None:
                self.halo.end(token)
                if ev:
                    ev[2].record()
                ops.stats_accumulate(s, ops.fd_lib_gram(V[cut:], D0, D1, DT, fold_of_frame=fof[cut:], **kw))
                if ev:
                    ev[3].record()
            if ev:
                iv = [(ev[0], ev[1])] + ([(ev[2], ev[3])] if token is not None else [])
                if record:
                    self.k1_ev.append(iv)
                self.pass_ev.append((ev[0], ev[3] if token is not None else ev[1]))
            return s

        def reduce(self, stats):
            if self.w == 1:
                return stats
            if self.comm is not None and self.kind == "flag":
                return self.comm.allreduce(stats)
            return slabs.allreduce_stats(stats)

        def step(self, library=L.LIB_KS_TRUE, block=BLOCK, variant=L.VARIANT_AUTO, record=False):
            stats = None
            for ps in range(self.passes):
                if self.passes > 1:
                    self.fill(ps)
                s = self.k1(ps, library, block, variant, record)
                stats = s if stats is None else ops.stats_accumulate(stats, s)
            if record and self.w > 1:
                e_a = torch.cuda.Event(enable_timing=True)
                e_a.record()
            stats = self.reduce(stats)
            p = L.LIB_WIDTH[library]
            self.last_stats = stats
            out = ops.stridge_batched(stats[0], p, dialect=L.STRIDGE_KS, flags=L.STRIDGE_RMS_PRESCALE, alphas=alphas,
                                      thresholds=thrs, max_iter=25, const_cols=[0] if p in (7, 9) else [],
                                      eval_stats=stats[1])
            if record and self.w > 1:
                e_b = torch.cuda.Event(enable_timing=True)
                e_b.record()
                self.tail_ev.append((e_a, e_b))
            return out

        def time(self, steps, warmup):
            """(ms per step: device time of the hot path, max over ranks; K1 ms per step list)"""
            for _ in range(warmup):
                out = self.step()
            barrier() if self.w > 1 else torch.cuda.synchronize()
            self.k1_ev, self.pass_ev, self.tail_ev = [], [], []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out = self.step(record=True)
            e1.record()
            barrier() if self.w > 1 else torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if self.passes > 1:
                # the refill between passes sits inside the bracketed region: count only the hot-path intervals (+ K3)
                ms = float(sum(a.elapsed_time(b) for a, b in self.pass_ev))
            ms = max_over_ranks(ms) if self.w > 1 else ms
            k1_all = [sum(a.elapsed_time(b) for a, b in iv) for iv in self.k1_ev]
            # this rank's mean time from the end of K1 to the end of K3 (all-reduce: waits for the slowest rank) and from
            # the start of the exchange to the start of K1 (barrier: waits for the slowest rank's previous step)
            self.phase_ms = None
            if self.w > 1 and self.tail_ev:
                tail = float(np.mean([a.elapsed_time(b) for a, b in self.tail_ev]))
                head = float(np.mean([p[0].elapsed_time(k[0][0]) for p, k in zip(self.pass_ev, self.k1_ev)])) if self.passes == 1 else None
                self.phase_ms = (head, tail)
            return ms / steps, k1_all, out

        def parity(self):
            """Hardware parity of the sharded path on THIS job's data (all ranks call it): the frame the exchange
            delivered is bit-identical to the true frame (the generator is deterministic in the global frame index),
            and the all-reduced statistics equal the rank-ordered sum of the per-slab statistics computed from slabs
            that hold the TRUE trailing frame."""
            if self.w == 1:
                return None
            ps = self.passes - 1
            a, b = self.sub[ps]
            self.fill(ps)                                   # trailing frame zeroed on ranks that must receive it
            self.step()
            torch.cuda.synchronize()
            got = self.last_stats.clone()
            true_frame = ops.synth_field(1, self.A, self.A, t_offset=b, T_total=1024, seed=0, noise=0.05)[0]
            frame_ok = bool(torch.equal(self.U[b - a], true_frame)) if not self.kind.endswith("means") else None
            self.U[b - a].copy_(true_frame)
            mine = ops.fd_lib_gram(self.view(ps), D0, D1, DT, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=BLOCK,
                                   n_folds=2, fold_of_frame=self.fof[ps])
            allp = [torch.empty_like(mine) for _ in range(self.w)]
            dist.all_gather(allp, mine)
            want = allp[0].clone()
            for q in range(1, self.w):
                want += allp[q]
            cs, ew = stats_rel_err(got.cpu().numpy(), want.cpu().numpy(), 3)
            flags = torch.tensor([1.0 if frame_ok in (True, None) else 0.0, 1.0 if torch.equal(got, want) else 0.0, -cs, -ew],
                                 dtype=torch.float64, device="cuda")
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)     # the worst rank decides
            f = flags.cpu().numpy()
            if self.kind != "flag" and self.r < self.w - 1:
                self.U[b - a].zero_()
            return {"halo_frame_bit_identical": (bool(f[0]) if frame_ok is not None else "n/a (block means travel, not the frame)"),
                    "allreduce_equals_rank_ordered_slab_sum_bitwise": bool(f[1]),
                    "allreduce_vs_slab_sum_max_rel": float(-f[2]), "allreduce_vs_slab_sum_max_rel_entrywise": float(-f[3])}

        def free(self):
            if self.comm is not None and self.kind == "flag":
                self.comm.release(self.U)
            self.U = None
            self.fof = None
            torch.cuda.empty_cache()

    def small_shard_parity():
        """Sharded == unsharded == oracle on a small stack, through every halo mechanism this run can set up (all ranks
        call it).  Each rank generates the WHOLE small stack locally, so the unsharded statistics are at hand."""
        from oracle import gram, ks2d as O

        A0s, A1s, tb_rank = 64, 256, 4
        g_rows_s = world * tb_rank * 3 + 2                      # ragged last t-block on the last rank
        whole = ops.synth_field(g_rows_s + 1, A0s, A1s, t_offset=0, T_total=64, seed=5, noise=0.05)
        split = int(0.7 * g_rows_s) // 3 * 3
        fof = (np.arange(g_rows_s) >= split).astype(np.int32)
        kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=BLOCK, n_folds=2)
        ref = ops.fd_lib_gram(whole, D0, D1, DT, fold_of_frame=fof, **kw)
        lo, hi = slabs.slab_bounds(g_rows_s, 3, world)[rank]
        res = {}
        modes = [("flag_symmetric_slab", "flag"), ("flag_published_frame", "flag_pub"), ("peer", "peer"), ("means", "means"),
                 ("send_recv", "send_recv"), ("send_recv_means", "send_recv_means")]
        for tag, mode in modes:
            if mode.startswith("flag") and comm is None:
                continue
            if mode == "flag":
                Ul = comm.slab((hi - lo + 1, A0s, A1s))
            else:
                Ul = torch.empty((hi - lo + 1, A0s, A1s), dtype=torch.float64, device="cuda")
            Ul.copy_(whole[lo:hi + 1])
            if rank < world - 1:
                Ul[-1].zero_()
            if mode.startswith("flag"):
                got = slabs.sharded_stats(Ul, D0, D1, DT, fold_of_frame=fof[lo:hi], comm=comm, **kw)
            else:
                with _StdoutToStderr():
                    ph = slabs.PeerHalo((A0s, A1s), peer_memory=not mode.startswith("send_recv"))
                got = slabs.sharded_stats(Ul, D0, D1, DT, fold_of_frame=fof[lo:hi], peer_halo=ph,
                                          block_means_halo=mode.endswith("means"), **kw)
            torch.cuda.synchronize()
            cs, ew = stats_rel_err(got.cpu().numpy(), ref.cpu().numpy(), 3)
            frame_ok = 1.0 if (mode.endswith("means") or rank == world - 1 or torch.equal(Ul[-1], whole[hi])) else 0.0
            f = torch.tensor([frame_ok, -cs, -ew], dtype=torch.float64, device="cuda")
            dist.all_reduce(f, op=dist.ReduceOp.MIN)
            f = f.cpu().numpy()
            res[tag] = {"halo_frame_bit_identical": bool(f[0]) if not mode.endswith("means") else "n/a",
                        "sharded_vs_unsharded_max_rel": float(-f[1]), "sharded_vs_unsharded_max_rel_entrywise": float(-f[2])}
            if mode == "flag":
                comm.release(Ul)
            del Ul
        # the unsharded statistics themselves against the oracle (rank 0)
        if rank == 0:
            wn = whole.cpu().numpy()
            nm, terms = O.build_dictionary_true(wn[:-1], D0, D1)
            X, y = O.build_blockwise_dataset((wn[1:] - wn[:-1]) / DT, terms, nm, block_t=3, block_x=8, block_y=8)
            tb = np.repeat(np.arange(-(-g_rows_s // 3)), (A0s // 8) * (A1s // 8))
            te = tb * 3 >= split
            oref = np.stack([gram.pack_stats(X[~te], y[~te]), gram.pack_stats(X[te], y[te])])
            cs, ew = stats_rel_err(ref.cpu().numpy(), oref, 3)
            res["unsharded_vs_oracle_max_rel"] = cs
            res["unsharded_vs_oracle_max_rel_entrywise"] = ew
        res["stack"] = f"{g_rows_s + 1}x{A0s}x{A1s}, {world} slabs, ragged last t-block"
        return res

    # ---- the contract workload (c4 per GPU, weak) or, with --workload c5, the strong-scaling stack as the main line
    if args.workload == "c5":
        A_main = C5_SIZE if args.size == 2048 else args.size
        g_rows = (C5_FRAMES if args.frames == 1024 else args.frames) - 1
        args.skip_e2e = args.skip_variants = args.skip_c5 = True
    else:
        A_main, g_rows = args.size, world * (args.frames - 1)
    if world > 1 and halo_kind != "flag":
        with _StdoutToStderr():
            halo = slabs.PeerHalo((A_main, A_main), peer_memory=not halo_kind.startswith("send_recv"))
    job = Job(A_main, g_rows, world, rank, comm, halo, halo_kind if world > 1 else "none")
    T, A = job.Tbuf, job.A

    # ---- parity gates before any timing
    parity = {}
    if rank == 0:
        from oracle import gram, ks2d as O

        small = job.U[:7, :64, :128].contiguous()
        s_gpu = ops.fd_lib_gram(small, D0, D1, DT, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=BLOCK).cpu().numpy()[0]
        sn = small.cpu().numpy()
        nm, terms = O.build_dictionary_true(sn[:-1], D0, D1)
        X, y = O.build_blockwise_dataset((sn[1:] - sn[:-1]) / DT, terms, nm, block_t=3, block_x=8, block_y=8)
        ref = gram.pack_stats(X, y)
        cs, ew = stats_rel_err(s_gpu, ref, 3)
        assert s_gpu[0] == ref[0] and cs <= 1e-10, f"parity gate failed: {cs}"
        parity["k1_vs_oracle_small"] = {"max_rel_cauchy_schwarz_scale": cs, "max_rel_entrywise": ew,
                                        "tolerance": 1e-10, "stack": "7x64x128 sub-stack of the workload"}
        if not args.no_cpu and args.frames >= 64 and A >= 1024:
            # the tiled kernel at the workload's own width against the oracle (25 frames x full width x 256 rows)
            sub = job.U[:25, :256, :].contiguous()
            s_gpu = ops.fd_lib_gram(sub, D0, D1, DT, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=BLOCK,
                                    variant=L.VARIANT_TILED).cpu().numpy()[0]
            sn = sub.cpu().numpy()
            nm, terms = O.build_dictionary_true(sn[:-1], D0, D1)
            X, y = O.build_blockwise_dataset((sn[1:] - sn[:-1]) / DT, terms, nm, block_t=3, block_x=8, block_y=8)
            cs, ew = stats_rel_err(s_gpu, gram.pack_stats(X, y), 3)
            assert cs <= 1e-10, f"parity gate (full width) failed: {cs}"
            parity["k1_tiled_vs_oracle_full_width"] = {"max_rel_cauchy_schwarz_scale": cs, "max_rel_entrywise": ew,
                                                       "stack": f"25x256x{A} sub-stack of the workload"}
    if world > 1:
        parity["mode"] = halo_kind + (f" (PeerComm unavailable: {comm_why})" if comm_why else "")
        parity["workload"] = job.parity()
        parity["small"] = small_shard_parity()
        bad = [k for k, v in parity["small"].items() if isinstance(v, dict) and
               (v["sharded_vs_unsharded_max_rel"] > 1e-12 or v["halo_frame_bit_identical"] is False)]
        w = parity["workload"]
        parity["green"] = bool(not bad and w["allreduce_vs_slab_sum_max_rel"] <= 1e-12 and w["halo_frame_bit_identical"] is not False)
        if rank == 0 and not parity["green"]:
            print("SHARDED PARITY FAILED:", json.dumps(parity), file=sys.stderr)
        job.fill(job.passes - 1)

    # ---- main timed region
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("PG_BENCH_NO_SAMPLER"):
        sampler.start()
    for _ in range(args.warmup):
        job.step()
    barrier()
    launches0 = lib.pg_launch_count()
    ms_step, k1_all, out = job.time(args.steps, 0)
    launches = lib.pg_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    k1_ms, k1_best = float(np.mean(k1_all)), float(np.min(k1_all))
    k1_ranks, phase_ranks = None, None
    if world > 1:
        # the step is the max over ranks: every rank's own K1 mean, so that the line shows the skew between the GPUs
        mine = torch.tensor([k1_ms], dtype=torch.float64, device="cuda")
        allk = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allk, mine)
        k1_ranks = [round(float(x.item()), 3) for x in allk]
        ph = getattr(job, "phase_ms", None)
        if ph is not None:
            mine = torch.tensor([ph[0] if ph[0] is not None else -1.0, ph[1]], dtype=torch.float64, device="cuda")
            allp = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allp, mine)
            phase_ranks = {"exchange_to_k1_start_ms": [round(float(x[0].item()), 3) for x in allp],
                           "k1_end_to_k3_end_ms": [round(float(x[1].item()), 3) for x in allp]}
    pts_k1 = sum(b - a + 1 for a, b in job.sub) * A * A       # points the K1 launches of one step read
    pts_step = (job.hi - job.lo + 1) * A * A                   # points this rank processes per step
    if args.workload == "c5":
        value = (g_rows + 1) * A * A / (ms_step * 1e-3)
    else:
        value = world * pts_step / (ms_step * 1e-3)

    best = int(out["best"].cpu()[0])
    coef = out["coef"].cpu().numpy()[0].reshape(-1, len(names))[best]

    # ---- end to end through the public API with pinned HOST buffers
    e2e_value, h2d, ms_e2e, e2e_frames, host_binding = None, 0, None, 0, None
    if not args.skip_e2e:
        cpus_before = os.sched_getaffinity(0)
        if not os.environ.get("PG_BENCH_NO_BIND"):
            # one process per GPU: keep this rank (and the pinned memory it first-touches below) on the GPU's NUMA node
            from pde_b200 import _xfer
            host_binding = _xfer.bind_host_to_gpu()
        e2e_frames = min(job.Tbuf, args.e2e_frames if args.e2e_frames > 0 else (job.Tbuf if world == 1 else 256))
        try:
            host = torch.empty((e2e_frames, A, A), dtype=torch.float64).pin_memory()
        except RuntimeError:                 # the box cannot pin that much: bounded sample
            e2e_frames = min(e2e_frames, 256)
            host = torch.empty((e2e_frames, A, A), dtype=torch.float64).pin_memory()
        for lo_ in range(0, e2e_frames, 64):
            host[lo_:lo_ + 64].copy_(job.U[lo_:min(e2e_frames, lo_ + 64)])
        torch.cuda.synchronize()
        bufs = [torch.empty((96 + 1, A, A), dtype=torch.float64, device="cuda") for _ in range(2)]
        fof_e = (np.arange(e2e_frames - 1) >= int(0.7 * (e2e_frames - 1)) // 3 * 3).astype(np.int32)
        res_host = torch.empty((30, len(names)), dtype=torch.float64).pin_memory()

        def e2e_step():
            st = slabs.fit_streamed(host, D0, D1, DT, library=L.LIB_KS_TRUE, block=BLOCK, fold_of_frame=fof_e, n_folds=2,
                                    slab_frames=96, buffers=bufs)
            if world > 1:
                st = comm.allreduce(st) if comm is not None else slabs.allreduce_stats(st)
            o = ops.stridge_batched(st[0], 3, dialect=L.STRIDGE_KS, flags=L.STRIDGE_RMS_PRESCALE, alphas=alphas,
                                    thresholds=thrs, max_iter=25, eval_stats=st[1])
            res_host.copy_(o["coef"].reshape(30, 3), non_blocking=False)

        e2e_steps = max(2, min(args.steps, 3))
        ms_e2e = timed(e2e_step, e2e_steps, 1)
        e2e_value = world * e2e_frames * A * A * e2e_steps / (ms_e2e * 1e-3)
        n_slabs = -(-(e2e_frames - 1) // 96)
        h2d = (e2e_frames + n_slabs - 1) * A * A * 8
        del host, bufs
        os.sched_setaffinity(0, cpus_before)                  # the CPU legs below use every core again

    # ---- variants (fewer steps): other libraries / estimators on the same stack
    variants = {}
    sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
    fp64_peak = 148 * 64 * sm_hz * 1e6

    FP64_PER_POINT = fp64_per_point()

    def variant_entry(name, ms, steps_):
        v = world * pts_k1 * steps_ / (ms * 1e-3)
        e = {"value": v, "unit": UNIT, "alg_GBps_per_gpu": 8 * v / world / 1e9,
             "roofline_hbm_frac": 8 * v / world / 1e9 / hbm_peak()[0]}
        if name in FP64_PER_POINT:
            rate = v / world * FP64_PER_POINT[name]
            e["roofline_fp64"] = {"fp64_instr_per_point": FP64_PER_POINT[name], "achieved_per_s": rate,
                                  "peak_per_s": fp64_peak, "frac": rate / fp64_peak,
                                  "peak": f"148 SMs x 64 fp64 lanes per clock x {sm_hz:.0f} MHz (median SM clock of the timed region)"}
        return e

    if not args.skip_variants:
        variants["true_p3_block388"] = variant_entry("true_p3_block388", k1_ms, 1)
        variants["true_p3_block388"]["note"] = "the contract line's K1 (k1_ms), for the fp64 roofline beside the HBM one"
        for name, kw in [("rich_p9_block388", dict(library=L.LIB_KS_RICH)),
                         ("true_adv_p5_block388", dict(library=L.LIB_KS_TRUE_ADV))]:
            ms = timed(lambda: job.step(**kw), 3, 1)
            variants[name] = variant_entry(name, ms, 3)

        # full-grid pointwise rows (every grid point a row; SURVEY 8d C4 i/ii): fp64-issue bound, not HBM bound
        def pointwise_step(dialect, library, p, sdialect, flags):
            s = ops.fd_lib_gram(job.view(0), D0, D1, DT, dialect=dialect, library=library, block=(1, 1, 1),
                                fold_of_frame=job.fof[0], n_folds=2)
            s = job.reduce(s)
            return ops.stridge_batched(s[0], p, dialect=sdialect, flags=flags, alphas=alphas, thresholds=thrs,
                                       max_iter=25 if sdialect == L.STRIDGE_KS else 10,
                                       eval_stats=s[1])

        job.fill(job.passes - 1, true_halo=True)    # these legs do no exchange: give the slab its true trailing frame
        for name, a in [("ks_true_p3_pointwise", (L.FD_KS_PERIODIC, L.LIB_KS_TRUE, 3, L.STRIDGE_KS, L.STRIDGE_RMS_PRESCALE)),
                        ("basic_usage_p6_pointwise", (L.FD_BASIC_TRIM, L.LIB_BASIC, 6, L.STRIDGE_BASIC, 0))]:
            ms = timed(lambda: pointwise_step(*a), 3, 1)
            variants[name] = variant_entry(name, ms, 3)
            variants[name]["bound"] = "fp64 issue (64 DFMA/clk/SM)"

    # ---- BASELINE configs[2]: patch-based spatial ensemble on a laser-image-shaped 1024x1024x500 float32 stack,
    # 8464 patches x (120 train + 40 test) sampled points: K2 (245-tap polynomial stencil rows) -> per-patch
    # statistics -> K3 (scikit-learn dialect).  Patches are embarrassingly parallel; every rank runs the same set.
    job_main_desc = dict(sub=job.sub, passes=job.passes, hi=job.hi, lo=job.lo)
    job.free()
    patch = None
    if not args.skip_variants and rank == 0:
        patch = patch_leg(args, torch, ops, L, alphas, thrs)

    # ---- BASELINE configs[0] / [1] (the reference's own CPU-runnable cases) through the drop-in API
    ref_cfgs = None
    if not args.skip_variants and rank == 0 and args.frames >= 1024:
        ref_cfgs = reference_configs_leg(args, torch, ops, K)

    # ---- BASELINE configs[4]: ONE 4096 x 4096 x 2048 stack, strong scaling, in the same run
    c5 = None
    if not args.skip_c5:
        g5 = C5_FRAMES - 1
        halo5 = None
        if world > 1 and halo_kind != "flag":
            with _StdoutToStderr():
                halo5 = slabs.PeerHalo((C5_SIZE, C5_SIZE), peer_memory=not halo_kind.startswith("send_recv"))
        j5 = Job(C5_SIZE, g5, world, rank, comm, halo5, halo_kind if world > 1 else "none")
        par5 = j5.parity()
        if world > 1:
            j5.fill(j5.passes - 1)
        steps5 = max(3, min(args.steps, 10))
        ms5, k1_5, out5 = j5.time(steps5, 3)
        b5 = int(out5["best"].cpu()[0])
        coef5 = out5["coef"].cpu().numpy()[0].reshape(-1, 3)[b5]
        pts5 = (g5 + 1) * C5_SIZE * C5_SIZE
        c5 = {"config": f"c5 (BASELINE configs[4]): ONE synthetic {C5_SIZE}x{C5_SIZE}x{C5_FRAMES} float64 stack cut into {world} time "
                        f"slab(s) of {j5.hi - j5.lo} row frames (whole t-blocks; the ragged last t-block stays with the last rank)"
                        + (f", streamed through one {j5.Tbuf}-frame buffer in {j5.passes} passes (generator refill excluded)" if j5.passes > 1 else "")
                        + ", same estimator as the contract line", "scaling": "strong",
              "ms_per_step": ms5, "value": pts5 / (ms5 * 1e-3), "unit": UNIT, "steps": steps5, "warmup": 3,
              "k1_ms": float(np.mean(k1_5)), "parity": par5,
              "selected": {"alpha": float(alphas.cpu()[b5 // 6]), "threshold": float(thrs.cpu()[b5 % 6]),
                           "coeffs": dict(zip(names, [float(c) for c in coef5]))}}
        j5.free()
        del j5
        if world > 1:
            # the single-GPU baseline of the SAME stack, timed by rank 0 alone in this run (the other ranks wait)
            barrier()
            if rank == 0:
                j1 = Job(C5_SIZE, g5, 1, 0)
                ms1, _, out1 = j1.time(max(2, min(args.steps, 5)), 2)
                b1 = int(out1["best"].cpu()[0])
                coef1 = out1["coef"].cpu().numpy()[0].reshape(-1, 3)[b1]
                c5["one_gpu_streamed_ms_per_step"] = ms1
                c5["one_gpu_passes"] = j1.passes
                c5["speedup_vs_1gpu_streamed"] = ms1 / ms5
                c5["same_selection_as_1gpu"] = bool(b1 == b5)
                c5["coeffs_max_rel_diff_vs_1gpu"] = float(np.max(np.abs(coef1 - coef5) / np.maximum(np.abs(coef1), 1e-300)))
                j1.free()
            barrier()
        else:
            c5["speedup_vs_1gpu_streamed"] = 1.0
            c5["one_gpu_streamed_ms_per_step"] = ms5

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peak, peak_src = hbm_peak()
    achieved = 8.0 * pts_k1 / (k1_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "k1_traffic.json"
    if tf.exists():
        tj = json.loads(tf.read_text())
        if tj.get("frames") and tj.get("size"):
            traffic = tj["dram_bytes_per_launch"] * pts_k1 / (tj["frames"] * tj["size"] ** 2)

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline_object()

    if args.workload == "c4":
        wl = workload_text(args, world)
    else:
        wl = (f"c5: ONE synthetic {A}x{A}x{g_rows + 1} float64 stack in {world} time slab(s) of {job_main_desc['hi'] - job_main_desc['lo']} row frames"
              f"{' streamed through one buffer in %d passes (generator refill excluded)' % job_main_desc['passes'] if job_main_desc['passes'] > 1 else ''}, "
              "KS periodic dialect, true dictionary p=3, block average (3,8,8), 2 time-holdout folds, 5x6 STRidge sweep")
    par_txt = "single GPU"
    if world > 1:
        par_txt = {"flag": f"time slabs x{world} in symmetric memory; 1-frame halo pulled by a copy engine over NVLink behind a flag the "
                           "single K1 launch polls; one-launch rank-ordered all-reduce of 2x18 doubles through peer memory (pg_comm_*)",
                   }.get(halo_kind, f"time slabs x{world}, 1-frame halo " + ("as 8x8 block means " if halo_kind.endswith("means") else "")
                         + f"({halo.mode if halo is not None else halo_kind}), NCCL all-reduce of 2x18 doubles")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if args.workload == "c5" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "l2": "inputs (%.1f GB per GPU) are larger than L2; no flush needed" % (pts_k1 * 8 / 1e9),
                   "parallelism": par_txt,
                   "selected": {"alpha": float(alphas.cpu()[best // 6]), "threshold": float(thrs.cpu()[best % 6]),
                                "coeffs": dict(zip(names, [float(c) for c in coef]))}},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k1_tiled_b88<KS_TRUE,2 folds>", "k1_ms": k1_ms,
                     "algorithmic_bytes": 8 * pts_k1, "peak_source": peak_src,
                     # fastest single step of the timed region (see profiles/README.md on the sustained-loop drift)
                     "k1_ms_best": k1_best, "frac_best": 8.0 * pts_k1 / (k1_best * 1e-3) / 1e9 / peak,
                     "k1_ms_steps": [round(x, 3) for x in k1_all], "k1_ms_ranks": k1_ranks, "phase_ms_ranks": phase_ranks,
                     "frac_of_nominal_8TBps": achieved / 8000.0},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 30 * 3 * 8,
                "h2d_GBps_per_gpu": (h2d * max(2, min(args.steps, 3)) / (ms_e2e * 1e-3) / 1e9) if ms_e2e else None,
                "sample": f"{e2e_frames} of {job_main_desc['hi'] - job_main_desc['lo'] + 1} frames per GPU streamed from pinned host memory in 96-frame slabs "
                          "(double-buffered), through pde_b200.slabs.fit_streamed + stridge_batched; bound by the host link "
                          "(PCIe Gen5 x16 per GPU; ranks that share a root complex share it)",
                "host_binding": host_binding},
        "gpu_launches": int(launches), "clocks": clocks, "parity": parity, "c5": c5, "variants": variants,
        "patch_ensemble": patch, "reference_configs": ref_cfgs,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def hbm_peak():
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        return float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def patch_leg(args, torch, ops, L, alphas, thrs):
    from pde_b200 import patch as PP

    torch.cuda.empty_cache()
    Tp, Hp = (500, 1024) if args.frames >= 1024 else (40, 256)
    U32 = ops.synth_field(Tp, Hp, Hp, seed=1, kind=1, noise=0.02).float()
    _, t_train, t_test = PP.time_split(Tp, 2, 0.7)
    coords = PP.patch_grid(Hp, Hp, 21, 10)
    tr_pts, te_pts = PP.sample_patch_points(np.random.default_rng(0), coords, Hp, Hp, 21, 3, t_train, t_test, 120)
    B = tr_pts.shape[0]
    W6 = ops._dev(PP.poly_stencil(2, 3, 3, 1.0, 0.1, 0.1))
    tr_d, te_d = ops._dev(tr_pts.reshape(-1, 3)), ops._dev(te_pts.reshape(-1, 3))
    a_d, t_d = ops._dev(np.array([0.01])), ops._dev(np.array([1e-5]))

    def patch_step():
        X, y = ops.poly_rows(U32, tr_d, W6, 2, 3, library=L.LIB_PATCH_FULL)
        Xt, yt = ops.poly_rows(U32, te_d, W6, 2, 3, library=L.LIB_PATCH_FULL)
        X, y, Xt, yt = X.view(B, 120, 8), y.view(B, 120), Xt.view(B, 40, 8), yt.view(B, 40)
        shift = X[:, 0, :].contiguous()
        st, mm = ops.rows_gram(X, y, shift=shift, want_minmax=True)
        se = ops.rows_gram(Xt, yt, shift=shift)
        return ops.stridge_batched(st[:, 0], 8, dialect=L.STRIDGE_SKLEARN, alphas=a_d, thresholds=t_d, max_iter=25,
                                   colminmax=mm[:, 0], shift=shift, eval_stats=se[:, 0])

    for _ in range(2):
        patch_step()
    torch.cuda.synchronize()
    # the pass is five short launches: replay it as one CUDA graph (fixed shapes; the library's scratch and
    # torch's buffers were allocated by the warm-up passes), eager launches if capture is not possible
    run_pass, mode = patch_step, "eager launches"
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            patch_step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):     # the stream the library's scratch was sized on
            graph_out = patch_step()
        graph.replay()
        torch.cuda.synchronize()
        ref_out = patch_step()
        if torch.equal(graph_out["coef"], ref_out["coef"]):
            run_pass, mode = graph.replay, "one CUDA graph replay per pass"
    except Exception as exc:  # pragma: no cover
        mode = f"eager launches (graph capture failed: {type(exc).__name__})"
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    # K3 alone over the whole ensemble with the reference's 5 x 6 (alpha, threshold) sweep: B x 30 fits per launch
    Xs, ys = ops.poly_rows(U32, tr_d, W6, 2, 3, library=L.LIB_PATCH_FULL)
    Xs, ys = Xs.view(B, 120, 8), ys.view(B, 120)
    sh = Xs[:, 0, :].contiguous()
    st_all, mm_all = ops.rows_gram(Xs, ys, shift=sh, want_minmax=True)
    k3 = lambda: ops.stridge_batched(st_all[:, 0], 8, dialect=L.STRIDGE_SKLEARN, alphas=alphas, thresholds=thrs,  # noqa: E731
                                     max_iter=25, colminmax=mm_all[:, 0], shift=sh)
    for _ in range(2):
        k3()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(5):
        k3()
    k1.record()
    torch.cuda.synchronize()
    ms_k3 = k0.elapsed_time(k1) / 5
    # the smoothing the script applies to the stack before the loop (patch:335, 343: gaussian_filter sigma 1.0, then 1.2)
    def prologue():
        return ops.gaussian_filter_frames(ops.gaussian_filter_frames(U32, 1.0), 1.2)

    prologue()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(3):
        prologue()
    g1.record()
    torch.cuda.synchronize()
    ms_gf = g0.elapsed_time(g1) / 3
    patch = {"workload": f"c3: {Hp}x{Hp}x{Tp} float32 stack, {B} patches x (120 train + 40 test) points, p=8, rt=2 rs=3 deg=3",
             "prologue": {"what": "gaussian_filter(sigma=1.0) then gaussian_filter(sigma=1.2) of every frame (patch:335,343), "
                                  "both axes fused per call (pg_reflect_gauss2d)",
                          "ms": ms_gf, "GBps_read_plus_write": 2 * 2 * 4.0 * Tp * Hp * Hp / (ms_gf * 1e-3) / 1e9},
             "k3_sweep_fits_per_s": B * 30 / (ms_k3 * 1e-3), "k3_sweep_ms": ms_k3,
             "stridge_fits_per_s": B / (ms * 1e-3), "stencil_points_per_s": B * 160 / (ms * 1e-3), "ms_per_pass": ms,
             "launch_mode": mode}
    if not args.no_cpu:
        from oracle import patch as OP
        from oracle import refload

        # parity gate + CPU timing on the first n_cpu patches of the SAME stack and the same RNG stream: the port with the
        # fixed stencil, and the unmodified reference loop body (per-point lstsq + scikit-learn) when it is staged
        n_cpu = 24
        Uh = U32.cpu().numpy()
        t0c = time.perf_counter()
        o = OP.run_patches(Uh, seed=0, max_patches=n_cpu)
        patch["cpu_port_fits_per_s"] = n_cpu / (time.perf_counter() - t0c)
        patch["cpu_port_sample"] = f"first {n_cpu} patches, fixed-stencil NumPy port, 1 core"
        mine = PP.fit_patches(U32, seed=0)
        Cg, Co = mine["C"][:n_cpu], o["C"]
        same = bool(np.array_equal(Cg != 0, Co != 0))
        nz = Co != 0
        rel = float(np.max(np.abs(Cg[nz] - Co[nz]) / np.abs(Co[nz]))) if nz.any() else 0.0
        patch["parity_vs_port"] = {"patches": n_cpu, "same_support": same, "coef_max_rel": rel, "tolerance": 1e-8}
        assert same and rel <= 1e-8, f"patch ensemble parity gate failed: support {same}, rel {rel}"
        try:                                   # the prologue against scipy itself on a few frames: time and bits
            from scipy.ndimage import gaussian_filter as sp_gf

            nf = 4
            t0c = time.perf_counter()
            ref_f = np.array([sp_gf(sp_gf(f, sigma=1.0), sigma=1.2) for f in Uh[:nf]])
            dt_sp = (time.perf_counter() - t0c) / nf
            got_f = ops.gaussian_filter_frames(ops.gaussian_filter_frames(U32[:nf].contiguous(), 1.0), 1.2).cpu().numpy()
            patch["prologue"].update(scipy_ms_per_frame=dt_sp * 1e3, scipy_frames=nf, bit_identical_to_scipy=bool(np.array_equal(got_f, ref_f)),
                                     speedup_vs_scipy_1core=dt_sp * 1e3 * Tp / ms_gf)
            assert patch["prologue"]["bit_identical_to_scipy"], "gaussian_filter prologue differs from scipy"
        except ImportError:
            patch["prologue"]["scipy"] = "not installed"
        if refload.available("patch"):
            pa = refload.load("patch")
            lib8 = pa.Library(names=list(PP.FULL_NAMES))
            n_ref = 4
            t0c = time.perf_counter()
            Cr = []
            for k in range(n_ref):
                pts = [tuple(int(v) for v in r) for r in mine["train_pts"][k]]
                Xr, yr = pa.build_dataset(Uh, pts, rt=2, rs=3, deg=3, dt=1.0, dx=0.1, dy=0.1, lib=lib8)
                Cr.append(pa.stridge(Xr, yr, alpha=0.01, threshold=1e-5))
            dt_ref = time.perf_counter() - t0c
            Cr = np.stack(Cr)
            nzr = Cr != 0
            patch["reference_fits_per_s"] = n_ref / dt_ref
            patch["parity_vs_reference"] = {
                "patches": n_ref, "same_support": bool(np.array_equal(mine["C"][:n_ref] != 0, nzr)),
                "coef_max_rel": float(np.max(np.abs(mine["C"][:n_ref][nzr] - Cr[nzr]) / np.abs(Cr[nzr]))) if nzr.any() else 0.0,
                "note": "unmodified build_dataset (per-point lstsq) + stridge (scikit-learn) of baseline/_ref; the lstsq-vs-stencil "
                        "floor is measured by tools/patch_floor.py"}
    return patch


def reference_configs_leg(args, torch, ops, K):
    """A HOST NumPy stack of the script's default shape (2000 x 100 x 100) goes in, the fitted model comes out (host ->
    device copy, kernels, result read-back all inside the timed call); beside it the CPU port of the same call and,
    when baseline/_ref is staged, the reference's own main() (cpu_baseline.kind "reference": its simulate() included,
    as the script has no entry point without it)."""
    Uc = ops.synth_field(2000, 100, 100, seed=2, noise=0.05).cpu().numpy()
    ref_cfgs = {"workload": "synthetic 2000x100x100 float64 HOST stack (the script's default grid), dx = dy = 0.5, DT = 1e-3; "
                            "wall clock of one fit_from_field call incl. host<->device copies"}
    cases = {"c1_pointwise_50k_true": dict(method="pointwise", dictionary="true"),
             "c2_blockwise388_true": dict(method="blockwise", dictionary="true"),
             "c2_blockwise388_rich_sweep": dict(method="blockwise", dictionary="rich", grid_search=True)}
    for cname, ckw in cases.items():
        K.fit_from_field(Uc, 0.5, 0.5, 1e-3, **ckw)
        torch.cuda.synchronize()
        t0c = time.perf_counter()
        for _ in range(3):
            mine = K.fit_from_field(Uc, 0.5, 0.5, 1e-3, **ckw)
        torch.cuda.synchronize()
        ref_cfgs[cname] = {"ours_ms": (time.perf_counter() - t0c) / 3 * 1e3}
        if not args.no_cpu:
            from oracle import ks2d as O

            t0c = time.perf_counter()
            theirs = O.run_config(Uc, 0.5, 0.5, 1e-3, **ckw)
            ref_cfgs[cname]["cpu_port_ms"] = (time.perf_counter() - t0c) * 1e3
            ref_cfgs[cname]["same_support"] = bool(np.array_equal(np.asarray(mine["coeffs"]) != 0,
                                                                  np.asarray(theirs["coeffs"]) != 0))
    if not args.no_cpu:
        ref_cfgs["cpu_port_note"] = "vectorised NumPy port (oracle/), 1 process"
        from oracle import refload

        if refload.available("ks2d"):
            import contextlib
            import io
            from unittest import mock

            ks = refload.load("ks2d")
            runs = {"c1_pointwise_50k_true": [],
                    "c2_blockwise388_true": ["--method", "blockwise", "--perturbation", "N2_noise", "--noise-rel", "0.05"]}
            for cname, argv in runs.items():
                buf = io.StringIO()
                t0c = time.perf_counter()
                with mock.patch.object(sys, "argv", ["ks2d_stridge_benchmark.py"] + argv), contextlib.redirect_stdout(buf):
                    ks.main()
                t_all = time.perf_counter() - t0c
                t0c = time.perf_counter()
                ks.simulate(ks.SimConfig())
                t_sim = time.perf_counter() - t0c
                ref_cfgs[cname]["reference_main_ms"] = t_all * 1e3
                ref_cfgs[cname]["reference_main_minus_simulate_ms"] = (t_all - t_sim) * 1e3
            ref_cfgs["cpu_baseline"] = {"kind": "reference", "cores": 1,
                                        "sample": "the unmodified ks2d_stridge_benchmark.main() of baseline/_ref on its own default "
                                                  "simulated 2000x100x100 stack (C1, C2); simulate() timed separately and subtracted"}
    return ref_cfgs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="frames per GPU (default: BASELINE configs[3])")
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4 (default): BASELINE configs[3] per GPU, weak scaling (+ the c5 object); c5: configs[4] as the main line")
    ap.add_argument("--e2e-frames", type=int, default=0, help="frames of the host-buffer leg (0 = the whole stack at N = 1, 256 per GPU at N > 1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    ap.add_argument("--skip-variants", action="store_true", help="profiling runs: skip the other libraries")
    ap.add_argument("--skip-c5", action="store_true", help="skip the strong-scaling c5 object")
    ap.add_argument("--port-only", action="store_true", help="reference arm: time only the NumPy port")
    args = ap.parse_args()
    # bounded CPU samples (frames, size); the environment overrides are the CPU tests' hook
    args.ref_sample = tuple(int(v) for v in os.environ.get("PG_BENCH_REF_SAMPLE", "49,512").split(","))
    args.port_sample = tuple(int(v) for v in os.environ.get("PG_BENCH_PORT_SAMPLE", "97,1024").split(","))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
