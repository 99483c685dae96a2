"""Why does K1 slow down inside a sustained loop?  

For each of several loops this prints, per step: the CUDA-event time, the EFFECTIVE SM clock measured inside the
kernel (cycle counter / %globaltimer of CTA 0: pg_fd_lib_gram's counters [4..7]) and the NVML readings taken while
it ran (requested SM clock, memory clock, power, temperature, throttle reasons).

    python tools/drift_probe.py [--frames 1024] [--size 2048] [--steps 40]

Loops: (a) K1 back to back; (b) K1 with a 20 ms host sleep between steps; (c) a device-to-device copy of the same
bytes back to back (torch copy_, read + write), whose per-step bandwidth shows what the memory path does by itself.
"""

import argparse
import json
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from pde_b200 import _lib as L  # noqa: E402
from pde_b200 import ops  # noqa: E402


class Nvml:
    def __init__(self):
        import pynvml

        pynvml.nvmlInit()
        self.n, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.rows, self.run = [], False

    def _loop(self):
        n = self.n
        while self.run:
            try:
                self.rows.append((time.perf_counter(), n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM),
                                  n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_MEM), n.nvmlDeviceGetPowerUsage(self.h) / 1e3,
                                  n.nvmlDeviceGetTemperature(self.h, n.NVML_TEMPERATURE_GPU),
                                  n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            except Exception:
                pass
            time.sleep(0.001)

    def __enter__(self):
        self.rows, self.run = [], True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.run = False
        self.t.join()

    def window(self, t0, t1):
        r = [x for x in self.rows if t0 <= x[0] <= t1]
        if not r:
            return None
        return dict(sm=int(np.median([x[1] for x in r])), mem=int(np.median([x[2] for x in r])),
                    w=round(float(np.max([x[3] for x in r]))), degc=int(np.max([x[4] for x in r])),
                    reasons=hex(int(np.bitwise_or.reduce([x[5] for x in r]))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--library", default="true", choices=["true", "rich"])
    args = ap.parse_args()
    T, A = args.frames, args.size
    U = ops.synth_field(T, A, A, seed=0, noise=0.05)
    fof = (np.arange(T - 1) >= int(0.7 * (T - 1)) // 3 * 3).astype(np.int32)
    fof_d = ops._dev(fof, torch.int32)
    lib = L.LIB_KS_TRUE if args.library == "true" else L.LIB_KS_RICH
    gb = 8.0 * T * A * A / 1e9

    def k1():
        return ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, dialect=L.FD_KS_PERIODIC, library=lib, block=(3, 8, 8),
                               fold_of_frame=fof_d, n_folds=2, return_nonfinite=True)

    for _ in range(3):
        k1()
    torch.cuda.synchronize()
    out = {"stack": f"{T}x{A}x{A}", "GB_per_step": gb, "library": args.library}

    def loop(tag, sleep_s):
        time.sleep(1.0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
        ctr, marks = [], []
        with Nvml() as nv:
            for k in range(args.steps):
                h0 = time.perf_counter()
                ev[2 * k].record()
                _, c = k1()
                ev[2 * k + 1].record()
                ctr.append(c)
                if sleep_s:
                    torch.cuda.synchronize()
                    marks.append((h0, time.perf_counter()))
                    time.sleep(sleep_s)
                else:
                    marks.append((h0, None))
            torch.cuda.synchronize()
            t_end = time.perf_counter()
        rows = []
        for k in range(args.steps):
            ms = ev[2 * k].elapsed_time(ev[2 * k + 1])
            c = ctr[k].cpu().numpy()
            eff = (c[6] - c[4]) / max(1, (c[7] - c[5])) * 1e3          # MHz
            rows.append(dict(step=k, ms=round(ms, 3), TBps=round(gb / ms, 3), sm_mhz_effective=round(float(eff), 1),
                             kernel_ms=round(float(c[7] - c[5]) / 1e6, 3)))
        # NVML windows: when steps are queued back to back the host runs ahead, so split the whole span evenly
        t0 = marks[0][0]
        for k in range(args.steps):
            if sleep_s:
                w = nv.window(marks[k][0], marks[k][1])
            else:
                span = (t_end - t0) / args.steps
                w = nv.window(t0 + k * span, t0 + (k + 1) * span)
            rows[k]["nvml"] = w
        out[tag] = rows
        print(tag, file=sys.stderr)
        for r in rows:
            print("  ", r, file=sys.stderr)

    loop("k1_back_to_back", 0.0)
    loop("k1_with_20ms_gaps", 0.02)

    # plain copy of the same number of bytes READ (so twice the traffic): what the memory path does alone
    n = T * A * A // 4
    a = torch.empty(n, dtype=torch.float64, device="cuda")
    b = torch.empty_like(a)
    del U
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    time.sleep(1.0)
    steps = args.steps * 2
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    with Nvml() as nv:
        t0 = time.perf_counter()
        ev[0].record()
        for k in range(steps):
            b.copy_(a)
            ev[k + 1].record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    rows = []
    for k in range(steps):
        ms = ev[k].elapsed_time(ev[k + 1])
        span = (t1 - t0) / steps
        rows.append(dict(step=k, ms=round(ms, 3), TBps_read_plus_write=round(2 * n * 8 / 1e9 / ms, 3),
                         nvml=nv.window(t0 + k * span, t0 + (k + 1) * span)))
    out["copy_back_to_back"] = rows
    print("copy_back_to_back", file=sys.stderr)
    for r in rows:
        print("  ", r, file=sys.stderr)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
