"""Turn the scratch ncu outputs under gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r02

Reads gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum launch list of bench.py),
gpurun_out/prof_k1_<tag>.ncu-rep (ncu --set full capture of the K1 kernel inside bench.py) and the quick_bench captures
prof_rich_<tag>, prof_adv_<tag>, prof_pw_ks_<tag>, prof_pw_basic_<tag> (.ncu-rep); writes
profiles/<tag>_launches_summary.csv, <tag>_launches_full.csv, <tag>_k1_tiled_ncu_summary.txt,
<tag>_k1_variants_ncu_summary.txt, profiles/k1_traffic.json (DRAM bytes per launch, read by bench.py for
roofline.traffic) and profiles/fp64_per_point.json (fp64 instructions per grid point of every K1 specialisation,
counted per opcode on the ncu source pages; read by bench.py for the fp64 roofline of the variants)."""

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
CMD = "python bench.py --steps 2 --warmup 3 --no-cpu --skip-e2e --skip-variants --skip-c5"
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
    rows = [r for r in csv.reader(open(SRC / f"launches_{tag}.csv")) if r and not r[0].startswith("==")]
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
    (OUT / f"{tag}_launches_full.csv").write_text((SRC / f"launches_{tag}.csv").read_text())
    print("\n".join(out[:10]))


def full(tag):
    raw = subprocess.run(["ncu", "-i", str(SRC / f"prof_k1_{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True,
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


def fp64_per_point(rep, points):
    """fp64 thread-instructions (DADD / DFMA / DMUL) per grid point and per-opcode shares, from the ncu source page."""
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    si = ti = wi = None
    fp = tot = warp_fp = warp_tot = 0.0
    sections = 0
    for r in rows:
        if r and r[0] == "Address":                      # one header per section (ncu prints one or more per captured launch)
            si, ti, wi = r.index("Source"), r.index("Thread Instructions Executed"), r.index("Instructions Executed")
            sections += 1
            continue
        if si is None or len(r) <= max(si, ti, wi):
            continue
        ops = [o for o in r[si].split() if not o.startswith("@")]
        if not ops:
            continue
        n, w = float(r[ti] or 0), float(r[wi] or 0)
        tot += n
        warp_tot += w
        if ops[0].split(".")[0] in ("DADD", "DFMA", "DMUL"):
            fp += n
            warp_fp += w
    points = points * max(sections, 1)                   # `points` = grid points of ONE launch
    return {"fp64_per_point": fp / points, "instr_per_point": tot / points, "fp64_share_of_warp_instr": warp_fp / max(warp_tot, 1.0)}


def variants(tag):
    """ncu --set full captures of the other K1 specialisations (tools/quick_bench.py --frames 256 --only <case>)."""
    alg = 8 * 256 * 2048 * 2048
    pts = 256 * 2048 * 2048
    lines = ["# ncu --set full --clock-control none --import-source on -k regex:k1_tiled -s 2 -c 1  "
             "python tools/quick_bench.py [--pointwise] --frames 256 --only <case>",
             f"# {tag}, B200, 2048x2048x256 fp64; algorithmic bytes per launch = {alg / 1e9:.2f} GB",
             "# pointwise kernels are bound by fp64 issue (64 DFMA/clk/SM), see sm__pipe_fp64_cycles_active"]
    counts = {}
    k1 = SRC / f"prof_k1_{tag}.ncu-rep"
    if k1.exists():
        counts["true_p3_block388"] = fp64_per_point(k1, 1024 * 2048 * 2048)
    for rep, what, key in ((f"prof_rich_{tag}.ncu-rep", "blockwise (3,8,8), KS rich library p = 9", "rich_p9_block388"),
                           (f"prof_adv_{tag}.ncu-rep", "blockwise (3,8,8), KS true + advection p = 5", "true_adv_p5_block388"),
                           (f"prof_pw_ks_{tag}.ncu-rep", "pointwise, KS dialect, true library p = 3", "ks_true_p3_pointwise"),
                           (f"prof_pw_basic_{tag}.ncu-rep", "pointwise, basic_usage dialect p = 6", "basic_usage_p6_pointwise")):
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
            g = lambda n: float(r[hdr.index(n)]) * UNIT[units[hdr.index(n)]]
            dram = g("dram__bytes_read.sum") + g("dram__bytes_write.sum")
            counts[key] = fp64_per_point(f, pts)
            lines.append(f"derived: {pts / (ms / 1e3) / 1e9:.1f} G points/s; algorithmic GB/s = {alg / 1e9 / (ms / 1e3):.1f}; "
                         f"DRAM traffic / algorithmic = {dram / alg:.4f}; fp64 instructions per point = {counts[key]['fp64_per_point']:.2f} "
                         f"of {counts[key]['instr_per_point']:.2f} thread instructions per point")
    (OUT / f"{tag}_k1_variants_ncu_summary.txt").write_text("\n".join(lines) + "\n")
    if counts:
        json.dump({"source": f"ncu source pages of the {tag} captures (tools/summarize_profiles.py): DADD + DFMA + DMUL thread "
                             "instructions / grid points of the launch", "kernels": counts},
                  open(OUT / "fp64_per_point.json", "w"), indent=1)
    print("\n".join(lines[-8:]))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    OUT.mkdir(exist_ok=True)
    launches(tag)
    full(tag)
    variants(tag)
