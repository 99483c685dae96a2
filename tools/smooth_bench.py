"""Timing of the denoising prologue kernels (developer tool): reflect-padded time moving average, periodic Gaussian
(two separable passes), scipy-style reflect Gaussian, and the two-stack fused path, on a synthetic stack.

    python tools/smooth_bench.py [--frames 256] [--size 2048]
"""

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from pde_b200 import _lib as L  # noqa: E402
from pde_b200 import ops  # noqa: E402


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--only-gaussian-filter", action="store_true")
    args = ap.parse_args()
    T, A = args.frames, args.size
    U = ops.synth_field(T, A, A, seed=0, noise=0.05)
    gb = 8.0 * T * A * A / 1e9
    rows = []
    if args.only_gaussian_filter:
        for dt in (torch.float64, torch.float32):
            V = U.to(dt)
            for sg, two in ((1.0, False), (1.2, False), (3.0, False), (1.0, True)):
                ms = timed(lambda: ops.gaussian_filter_frames(V, sg, two_pass=two))
                gbv = gb * (1.0 if dt == torch.float64 else 0.5)
                print(json.dumps(dict(op=f"gaussian_filter_frames(sigma={sg}, {'two passes' if two else 'fused'}, {str(dt)[6:]})",
                                      ms=round(ms, 3), GBps_one_read_one_write=round(2 * gbv / ms * 1e3, 1))))
        return
    for w in (3, 5):
        ms = timed(lambda: ops.time_moving_average(U, w))
        rows.append(dict(op=f"time_moving_average(window={w})", ms=round(ms, 3), GBps_read_plus_write=round(2 * gb / ms * 1e3, 1)))
    # sigma < 2.82 px keeps ALL n taps per axis (the spectral cut-off makes the periodic Gaussian ring): only on small frames
    for sg in ((1.0, 3.0) if A <= 512 else (3.0,)):
        ms = timed(lambda: ops.gaussian_smooth_periodic(U, sg), iters=1)
        rows.append(dict(op=f"gaussian_smooth_periodic(sigma={sg}) [two passes, {len(ops.periodic_gaussian_taps(A, sg)[0])} taps]",
                         ms=round(ms, 3), GBps_read_plus_write=round(4 * gb / ms * 1e3, 1)))
    ms = timed(lambda: ops.gaussian_filter_frames(U, 1.0))
    rows.append(dict(op="gaussian_filter_frames(sigma=1) [scipy reflect, both axes fused]", ms=round(ms, 3),
                     GBps_read_plus_write=round(2 * gb / ms * 1e3, 1)))
    V = ops.time_moving_average(U, 3)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8))
    ms = timed(lambda: ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, Uy=V, **kw))
    rows.append(dict(op="fd_lib_gram two stacks (tiled)", ms=round(ms, 3), alg_GBps_one_stack=round(gb / ms * 1e3, 1)))
    ms = timed(lambda: ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, Uy=V, variant=L.VARIANT_GENERIC, **kw), iters=1)
    rows.append(dict(op="fd_lib_gram two stacks (generic)", ms=round(ms, 3), alg_GBps_one_stack=round(gb / ms * 1e3, 1)))
    ms = timed(lambda: ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, **kw))
    rows.append(dict(op="fd_lib_gram one stack (tiled)", ms=round(ms, 3), alg_GBps_one_stack=round(gb / ms * 1e3, 1)))
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
