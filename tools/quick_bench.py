"""Quick K1 timing probe (developer tool, not the contract bench): times pg_fd_lib_gram variants
on a synthetic stack with CUDA events and prints achieved algorithmic GB/s (8 B per grid point)."""

import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from pde_b200 import _lib as L  # noqa: E402
from pde_b200 import ops  # noqa: E402


class Nvml:
    """SM clock / power sampler (pynvml, ~2 ms period) for the timed loops."""

    def __init__(self):
        import threading

        import pynvml

        pynvml.nvmlInit()
        self.n, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.rows, self.run = [], False
        self.t = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        import time

        while self.run:
            try:
                self.rows.append((self.n.nvmlDeviceGetClockInfo(self.h, self.n.NVML_CLOCK_SM),
                                  self.n.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  self.n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        self.run = True
        self.t.start()
        return self

    def __exit__(self, *a):
        self.run = False
        self.t.join()

    def summary(self):
        import statistics

        if not self.rows:
            return {}
        return dict(sm_mhz_median=statistics.median(r[0] for r in self.rows), sm_mhz_min=min(r[0] for r in self.rows),
                    power_w_max=round(max(r[1] for r in self.rows), 1), reasons=hex(max(r[2] for r in self.rows)),
                    samples=len(self.rows))


def time_call(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for k in range(iters):
        fn()
        ev[k + 1].record()
    torch.cuda.synchronize()
    ts = [ev[k].elapsed_time(ev[k + 1]) for k in range(iters)]
    return min(ts), sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--width", type=int, default=None, help="columns (default: --size)")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--generic", action="store_true")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--pointwise", action="store_true", help="time the tiled pointwise kernel (every point a row)")
    ap.add_argument("--only", default=None, help="run a single named case")
    args = ap.parse_args()
    T, A = args.frames, args.size
    A1 = args.width or A
    U = ops.synth_field(T, A, A1, seed=0, noise=0.05)
    torch.cuda.synchronize()
    pts = T * A * A1
    fof = (torch.arange(T - 1) >= int(0.7 * (T - 1))).to(torch.int32).cuda()  # on the device: no per-call copy
    out = []
    cases = [("true_b388_tiled", L.LIB_KS_TRUE, (3, 8, 8), L.VARIANT_TILED, 1),
             ("true_b388_tiled_2folds", L.LIB_KS_TRUE, (3, 8, 8), L.VARIANT_TILED, 2),
             ("rich_b388_tiled_2folds", L.LIB_KS_RICH, (3, 8, 8), L.VARIANT_TILED, 2),
             ("adv_b388_tiled", L.LIB_KS_TRUE_ADV, (3, 8, 8), L.VARIANT_TILED, 1)]
    if args.pointwise:
        cases = [("ks_true_pointwise_tiled", L.LIB_KS_TRUE, (1, 1, 1), L.VARIANT_TILED, 1),
                 ("ks_true_pointwise_tiled_2folds", L.LIB_KS_TRUE, (1, 1, 1), L.VARIANT_TILED, 2),
                 ("ks_adv_pointwise_tiled", L.LIB_KS_TRUE_ADV, (1, 1, 1), L.VARIANT_TILED, 1),
                 ("ks_richnoadv_pointwise_tiled", L.LIB_KS_RICH_NOADV, (1, 1, 1), L.VARIANT_TILED, 1),
                 ("ks_rich_pointwise_tiled", L.LIB_KS_RICH, (1, 1, 1), L.VARIANT_TILED, 1),
                 ("basic_pointwise_tiled", L.LIB_BASIC, (1, 1, 1), L.VARIANT_TILED, 1),
                 ("basic_pointwise_tiled_2folds", L.LIB_BASIC, (1, 1, 1), L.VARIANT_TILED, 2)]
    if args.generic:
        cases += [("true_b388_generic", L.LIB_KS_TRUE, (3, 8, 8), L.VARIANT_GENERIC, 1),
                  ("true_pointwise_generic", L.LIB_KS_TRUE, (1, 1, 1), L.VARIANT_GENERIC, 1),
                  ("rich_pointwise_generic", L.LIB_KS_RICH, (1, 1, 1), L.VARIANT_GENERIC, 1)]
    for name, lib, block, variant, nf in cases:
        if args.only and name != args.only:
            continue
        kw = dict(dialect=L.FD_BASIC_TRIM if lib == L.LIB_BASIC else L.FD_KS_PERIODIC, library=lib, block=block,
                  variant=variant, n_folds=nf)
        if nf == 2:
            kw["fold_of_frame"] = fof
        with Nvml() as nv:
            best, avg = time_call(lambda: ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, **kw), iters=args.iters)
        rec = dict(case=name, ms_best=round(best, 3), ms_avg=round(avg, 3), gpts_per_s=round(pts / best / 1e6, 2),
                   alg_GBps=round(8 * pts / best / 1e6, 1), clocks=nv.summary())
        print(json.dumps(rec), flush=True)
        out.append(rec)
    if args.sweep:
        # TMA L2 promotion of the field's tensor map, read by the launcher at every call
        for promo in (0, 1, 2, 3):
            os.environ["PG_TMA_L2PROMO"] = str(promo)
            for name, lib, nf in (("true", L.LIB_KS_TRUE, 2), ("rich", L.LIB_KS_RICH, 2), ("true1f", L.LIB_KS_TRUE, 1)):
                kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=(3, 8, 8), variant=L.VARIANT_TILED, n_folds=nf)
                if nf == 2:
                    kw["fold_of_frame"] = fof
                best, avg = time_call(lambda: ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, **kw), iters=args.iters)
                print(json.dumps(dict(case=f"{name}_promo{promo}", ms_best=round(best, 3), ms_avg=round(avg, 3),
                                      alg_GBps=round(8 * pts / best / 1e6, 1))), flush=True)
    # plain device copy of the same bytes for scale (read + write)
    V = torch.empty_like(U[: T // 2])
    best, _ = time_call(lambda: V.copy_(U[: T // 2]), iters=args.iters)
    print(json.dumps(dict(case="torch_copy_half", ms_best=round(best, 3),
                          GBps_rw=round(2 * 8 * V.numel() / best / 1e6, 1))))


if __name__ == "__main__":
    main()
