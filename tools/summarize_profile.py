#!/usr/bin/env python
"""Turn an `ncu --set full` capture of bench.py into the small tracked summaries under
profiles/: a JSON with the numbers bench.py and DESIGN.md quote, and a text report with
the per-function instruction attribution (tools/ncu_lines.py).

usage: summarize_profile.py <report.ncu-rep> <round-tag> [groups] [ticks_per_launch] [kernel-substring (default step_kernel)]
"""
import contextlib
import csv
import io
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__icc_request_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    groups = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
    ticks = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    ksub = sys.argv[5] if len(sys.argv) > 5 else "step_kernel"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    kcol = hdr.index("Kernel Name")
    row = [r for r in body if ksub in r[kcol]][0]
    m = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            v = float(row[i].replace(",", ""))
            m[w] = v * SCALE.get(units[i], 1.0) if units[i] in SCALE else v
    out = {
        "round": tag, "kernel": row[kcol], "groups": groups, "replicas": 5, "ticks_per_launch": ticks,
        "command": f"ncu --set full --clock-control none --import-source on -k regex:{ksub} -s <warm-up launches> -c 1 "
                   "python bench.py --no-e2e --no-cpu --no-others --no-parity --steps 4 --warmup 3",
        "note": "under ncu the kernel runs with cold caches and serialised replays; use shares, not absolutes",
        "duration_s": m.get("gpu__time_duration.sum"),
        "dram_bytes_read": m.get("dram__bytes_read.sum"), "dram_bytes_write": m.get("dram__bytes_write.sum"),
        "dram_bytes_per_launch": (m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)),
        "metrics": m,
    }
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_{ksub}.json"), "w") as f:
        json.dump(out, f, indent=1)
    latest = "step_kernel_latest.json" if ksub == "step_kernel" else "dominant_kernel_latest.json"
    with open(os.path.join(ROOT, "profiles", latest), "w") as f:     # bench.py's roofline.traffic reads dominant_kernel_latest.json first
        json.dump(out, f, indent=1)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        print(f"# {tag}: per-function warp-instructions of {ksub}<5>, per 32 groups and tick "
              f"({groups // 32} x {ticks} ticks)")
        ncu_lines.by_function(rep, os.path.join(ROOT, "josefine_b200/csrc/libjosefine_b200.so"), ksub + "ILi5E",
                              os.path.join(ROOT, "josefine_b200/csrc/sym_fold.cuh" if ksub.startswith("sym") else "josefine_b200/csrc/raft_device.cuh"),
                              0, (groups // 32) * ticks)
    with open(os.path.join(ROOT, "profiles", f"{tag}_{ksub}_functions.txt"), "w") as f:
        f.write(buf.getvalue())
    print(json.dumps({k: out[k] for k in ("duration_s", "dram_bytes_read", "dram_bytes_write")}))


if __name__ == "__main__":
    main()
