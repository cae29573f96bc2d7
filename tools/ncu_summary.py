#!/usr/bin/env python
"""Turn the ncu artefacts a GPU call left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summary.py <tag>            # e.g. r01a

  gpurun_out/launches.csv      -> profiles/<tag>_launches.csv (copy) + per-kernel shares in the .md
  gpurun_out/prof_*.ncu-rep    -> profiles/<tag>_<name>.json  (key raw metrics per captured launch)
                                  profiles/<tag>_summary.md
The flow_iter capture also refreshes profiles/flow_iter_traffic.json (read by bench.py for
roofline.traffic).
"""
from __future__ import annotations

import csv
import glob
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0,
              "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def raw_rows(rep):
    # a capture may travel as its raw-page CSV (exported on the GPU box: gpurun returns at most 64 MiB) instead of the report
    if rep.endswith("_raw.csv"):
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                v *= UNIT_SCALE.get(units[i], 1.0)
                d[k] = v
        out.append(d)
    return out


def launch_shares(path):
    by = {}
    total = 0.0
    with open(path) as f:
        rows = [r for r in csv.reader(f) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "")
        v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
        n, t = by.get(name, (0, 0.0))
        by[name] = (n + 1, t + v)
        total += v
    return by, total


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    note = os.environ.get("OFC_PROFILE_NOTE")
    md = [f"# ncu summary `{tag}`", ""]
    if note:
        md += [note, ""]
    else:
        md += ["Captured on a B200 through `gpurun` with `tools/gpu_profile.sh` (command: `python tools/profile_step.py`,",
               "3 steps of the 1080p pipeline, 9 frames = 8 pairs per step).  ncu times are cold-cache and serialised:",
               "compare SHARES with bench.py's CUDA-event breakdown, not absolutes.", ""]
    lc = os.path.join(OUT, "launches.csv")
    if os.path.exists(lc):
        shutil.copy(lc, os.path.join(PROF, f"{tag}_launches.csv"))
        by, total = launch_shares(lc)
        md += ["## launch list (`--metrics gpu__time_duration.sum --clock-control none`)", "",
               "| kernel | launches | total us | share |", "|---|---|---|---|"]
        for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            md.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / total:.1f} % |")
        md.append("")
    reps = sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep")) + glob.glob(os.path.join(OUT, f"{tag}_*.ncu-rep")))
    have = {os.path.basename(r)[:-8] for r in reps}
    reps += [c for c in sorted(glob.glob(os.path.join(OUT, f"{tag}_*_raw.csv"))) if os.path.basename(c)[:-8] not in have]
    for rep in reps:
        base = os.path.basename(rep)
        name = base[5:-8] if base.startswith("prof_") else base[len(tag) + 1:-8]
        rows = raw_rows(rep)
        with open(os.path.join(PROF, f"{tag}_{name}.json"), "w") as f:
            json.dump(rows, f, indent=1)
        md += [f"## `{name}` (`ncu --set full --clock-control none --import-source on`)", ""]
        for d in rows:
            dur = d.get("gpu__time_duration.sum", 0.0)
            rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
            md += [f"- `{d['kernel']}` grid {d['grid']} block {d['block']}: {dur * 1e6:.1f} us, DRAM read {rd / 1e6:.1f} MB + "
                   f"write {wr / 1e6:.1f} MB ({(rd + wr) / max(dur, 1e-12) / 1e9:.0f} GB/s, "
                   f"{d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):.1f} % of DRAM peak), "
                   f"SM throughput {d.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):.1f} %, "
                   f"warps active {d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.1f} %, "
                   f"{int(d.get('launch__registers_per_thread', 0))} regs/thread, "
                   f"{d.get('smsp__inst_executed.sum', 0) / 1e6:.1f} M warp instructions, "
                   f"L2 hit {d.get('lts__t_sector_hit_rate.pct', 0):.1f} %"
                   + (f", tensor pipe active {d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']:.1f} %"
                      if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' in d else "")]
        md.append("")
        iters = [r for r in rows if "flow_iter_tmem" in r["kernel"]]
        if iters:
            d = max(iters, key=lambda r: r.get("gpu__time_duration.sum", 0.0))     # the full-resolution launch
            with open(os.path.join(PROF, f"{tag[:3]}_flow_iter_traffic.json"), "w") as f:
                json.dump({"source": f"profiles/{tag}_{name}.json", "kernel": d["kernel"], "grid": d["grid"],
                           "pairs_per_launch": int(os.environ.get("OFC_CHUNK", "9")) - 1,
                           "dram_bytes_per_launch": d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0),
                           "duration_us_under_ncu": d.get("gpu__time_duration.sum", 0.0) * 1e6}, f, indent=1)
    with open(os.path.join(PROF, f"{tag}_summary.md"), "w") as f:
        f.write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
