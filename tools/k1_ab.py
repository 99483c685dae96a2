"""Compact A/B probe of the blockwise K1 kernel under different experiment switches.

For every configuration (a set of environment variables, read by the library at each launch) this runs the C4
stack back to back and prints ONE line: the mean time of the first steps (SM clock still at its maximum), of the
last steps (after the box has pulled the clock down), the effective SM clock measured inside the kernel over those
last steps and time x clock (SM cycles per launch, the clock-independent cost on the SM side).

    python tools/k1_ab.py "PG_TILED_WS=0" "PG_TILED_WS=1" "PG_TILED_WS=1 PG_TILED_LEAD=-1" [--library rich] [--steps 40]
"""

import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from pde_b200 import _lib as L  # noqa: E402
from pde_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--library", default="true", choices=["true", "rich", "adv"])
    ap.add_argument("--rowfolds", action="store_true", help="two per-row folds instead of time folds")
    args = ap.parse_args()
    T, A = args.frames, args.size
    U = ops.synth_field(T, A, A, seed=0, noise=0.05)
    lib = {"true": L.LIB_KS_TRUE, "rich": L.LIB_KS_RICH, "adv": L.LIB_KS_TRUE_ADV}[args.library]
    kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=(3, 8, 8), n_folds=2, return_nonfinite=True)
    if args.rowfolds:
        nb = ((T - 1 + 2) // 3) * (A // 8) * (A // 8)
        kw["fold_of_row"] = ops._dev((np.random.default_rng(0).random(nb) < 0.3).astype(np.uint8), torch.uint8)
    else:
        fof = (np.arange(T - 1) >= int(0.7 * (T - 1)) // 3 * 3).astype(np.int32)
        kw["fold_of_frame"] = ops._dev(fof, torch.int32)
    gb = 8.0 * T * A * A / 1e9
    ref = None
    for cfg in args.configs:
        env = dict(kv.split("=", 1) for kv in cfg.split())
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            for _ in range(2):
                s, _c = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, **kw)
            torch.cuda.synchronize()
            time.sleep(1.5)      # let the clock recover from the previous configuration
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
            ctr = []
            for k in range(args.steps):
                ev[2 * k].record()
                s, c = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, **kw)
                ev[2 * k + 1].record()
                ctr.append(c)
            torch.cuda.synchronize()
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        ms = np.array([ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)])
        cs = [c.cpu().numpy() for c in ctr]
        mhz = np.array([(c[6] - c[4]) / max(1, (c[7] - c[5])) * 1e3 for c in cs])
        stats = s.cpu().numpy()
        if ref is None:
            ref = stats
        dev = float(np.max(np.abs(stats - ref) / (np.abs(ref) + 1e-300)))
        n_tail = max(5, args.steps // 3)
        print(json.dumps(dict(cfg=cfg, ms_first5=round(float(ms[1:6].mean()), 3), ms_min=round(float(ms.min()), 3),
                              ms_tail=round(float(ms[-n_tail:].mean()), 3), mhz_tail=round(float(mhz[-n_tail:].mean())),
                              mcycles_tail=round(float((ms[-n_tail:] * mhz[-n_tail:]).mean()), 1),
                              TBps_first5=round(gb / float(ms[1:6].mean()), 3), TBps_tail=round(gb / float(ms[-n_tail:].mean()), 3),
                              ms_mean=round(float(ms.mean()), 3), max_rel_dev_vs_first_cfg=dev)), flush=True)


if __name__ == "__main__":
    main()
