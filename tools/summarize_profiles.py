"""Turn the scratch ncu outputs under gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01

Reads gpurun_out/launches_r1.csv (ncu --metrics gpu__time_duration.sum launch list of bench.py) and
gpurun_out/prof_k1_r1.ncu-rep (ncu --set full capture of the K1 kernel), writes
profiles/<tag>_launches_summary.csv, <tag>_launches_full.csv, <tag>_k1_tiled_ncu_summary.txt and
profiles/k1_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic)."""

import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
SRC = ROOT / "gpurun_out"
CMD = "python bench.py --steps 2 --warmup 3 --no-cpu --skip-e2e --skip-variants"
ALG = 8 * 1024 * 2048 * 2048

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sector_op_read_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
TIME = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}


def launches(tag):
    rows = [r for r in csv.reader(open(SRC / "launches_r1.csv")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])[:80]
        us = float(r[vi].replace(",", "")) * TIME.get(r[ui], 1e-6) * 1e3
        tot[name] += us
        cnt[name] += 1
    T = sum(tot.values())
    timed = {k: v for k, v in tot.items() if "synth" not in k}
    Tt = sum(timed.values())
    out = [f"# ncu launch list summary ({tag}) -- command: {CMD}",
           "# ncu --metrics gpu__time_duration.sum --clock-control none -c 400   (per-launch times are cold-cache and serialised: compare SHARES)",
           f"# total device time over {sum(cnt.values())} launches: {T / 1000:.2f} ms; synth_kernel is the input generator (outside the timed region)",
           "kernel,launches,total_us,share_all,share_of_step_kernels"]
    for k, v in tot.most_common():
        out.append(f"\"{k}\",{cnt[k]},{v:.1f},{v / T:.4f},{(v / Tt if k in timed else 0):.4f}")
    (OUT / f"{tag}_launches_summary.csv").write_text("\n".join(out) + "\n")
    (OUT / f"{tag}_launches_full.csv").write_text((SRC / "launches_r1.csv").read_text())
    print("\n".join(out[:10]))


def full(tag):
    raw = subprocess.run(["ncu", "-i", str(SRC / "prof_k1_r1.ncu-rep"), "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none --import-source on -k regex:k1_tiled -s 3 -c 2  {CMD}",
             f"# {tag}, B200, workload c4 (2048x2048x1024 fp64, true dictionary, (3,8,8) blocks, 2 time-holdout folds)",
             f"# algorithmic bytes per launch = 8 B x 1024 x 2048 x 2048 = {ALG / 1e9:.2f} GB"]
    tr = []
    for r in rows[2:]:
        lines.append("----")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"{w} = {r[i]} {units[i]}")
        g = lambda n: float(r[hdr.index(n)]) * UNIT[units[hdr.index(n)]]
        tr.append(g("dram__bytes_read.sum") + g("dram__bytes_write.sum"))
        ms = float(r[hdr.index("gpu__time_duration.sum")]) * TIME[units[hdr.index("gpu__time_duration.sum")]]
        l2 = float(r[hdr.index("lts__t_sectors_srcunit_tex_op_read.sum")]) * 32
        lines.append(f"derived: algorithmic GB/s = {ALG / 1e9 / (ms / 1e3):.1f}; DRAM traffic / algorithmic = {tr[-1] / ALG:.4f}; "
                     f"L2->SM read requests / algorithmic = {l2 / ALG:.4f}")
    (OUT / f"{tag}_k1_tiled_ncu_summary.txt").write_text("\n".join(lines) + "\n")
    json.dump({"kernel": "k1_tiled_b88<KS_TRUE,2 folds,8 warps>", "frames": 1024, "size": 2048,
               "dram_bytes_per_launch": sum(tr) / len(tr),
               "source": f"profiles/{tag}_k1_tiled_ncu_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)"},
              open(OUT / "k1_traffic.json", "w"), indent=1)
    print("\n".join(lines[-28:]))


def pointwise(tag):
    """ncu --set full captures of the tiled pointwise kernel (tools/quick_bench.py --pointwise --frames 256)."""
    alg = 8 * 256 * 2048 * 2048
    lines = ["# ncu --set full --clock-control none --import-source on -k regex:k1_tiled_pw -s 2 -c 1  "
             "python tools/quick_bench.py --pointwise --frames 256 --only <case>",
             f"# {tag}, B200, 2048x2048x256 fp64, every grid point a row; algorithmic bytes per launch = {alg / 1e9:.2f} GB",
             "# bound: fp64 issue (64 DFMA/clk/SM), see sm__pipe_fp64_cycles_active"]
    for rep, what in (("prof_pw_ks.ncu-rep", "KS dialect, true library p = 3"), ("prof_pw_basic.ncu-rep", "basic_usage dialect p = 6")):
        f = SRC / rep
        if not f.exists():
            continue
        raw = subprocess.run(["ncu", "-i", str(f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            lines.append(f"---- {what}")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    lines.append(f"{w} = {r[i]} {units[i]}")
            ms = float(r[hdr.index("gpu__time_duration.sum")]) * TIME[units[hdr.index("gpu__time_duration.sum")]]
            pts = 256 * 2048 * 2048
            lines.append(f"derived: {pts / (ms / 1e3) / 1e9:.1f} G points/s; algorithmic GB/s = {alg / 1e9 / (ms / 1e3):.1f}")
    (OUT / f"{tag}_k1_pointwise_ncu_summary.txt").write_text("\n".join(lines) + "\n")
    print("\n".join(lines[-8:]))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    OUT.mkdir(exist_ok=True)
    launches(tag)
    full(tag)
    pointwise(tag)
