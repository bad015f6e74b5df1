#!/usr/bin/env python
"""Per-source-line attribution of an ncu capture.

ncu's CLI source page lists SASS with per-instruction counters but without the
CUDA-C line; nvdisasm lists the same SASS with `//## File ... line N` markers
(needs -lineinfo).  Both list one function's instructions in the same order, so
joining by index gives executed-instruction and stall-sample counts per source
line -- including lines inlined from raft_device.cuh.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [capture-index]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def disasm_lines(so_path, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=tmp,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout
    out, active, pending = [], False, None
    for line in txt.splitlines():
        if line.startswith(".text."):
            active = kernel_sub in line
            pending = None
            continue
        if not active:
            continue
        if line.startswith("\t.section") or line.startswith(".section"):
            active = False
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            if pending is None:  # first marker after an instruction = innermost frame
                pending = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), pending or (out[-1][2] if out else ("?", 0))))
            pending = None
    return out


def ncu_sass(rep, kernel_sub, which):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    starts = [i for i in starts if kernel_sub in rows[i][1]]
    s = starts[which]
    ends = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and i > s]
    e = ends[0] if ends else len(rows)
    hdr = rows[s + 1]
    return hdr, rows[s + 2:e]


def main():
    rep, so, ksub = sys.argv[1:4]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    dis = disasm_lines(so, ksub)
    hdr, body = ncu_sass(rep, ksub.replace("ILi", "<").split("<")[0], which)
    col = {h: i for i, h in enumerate(hdr)}
    if len(dis) != len(body):
        print(f"# WARNING: nvdisasm has {len(dis)} instructions, ncu has {len(body)}; joining the common prefix")
    n = min(len(dis), len(body))
    ex, samp, longsb = collections.Counter(), collections.Counter(), collections.Counter()
    for i in range(n):
        key = dis[i][2]
        r = body[i]
        ex[key] += int(r[col["Instructions Executed"]] or 0)
        samp[key] += int(r[col["# Samples"]] or 0)
        longsb[key] += int(r[col.get("stall_long_sb", col["# Samples"])] or 0)
    tot_e, tot_s = sum(ex.values()), sum(samp.values())
    print(f"# {n} SASS instructions, {tot_e} warp-instructions executed, {tot_s} stall samples")
    print(f"{'file:line':34s} {'executed':>10s} {'%':>6s} {'samples':>8s} {'%':>6s} {'long_sb':>8s}")
    for key, v in sorted(ex.items(), key=lambda kv: -(kv[1] + 50 * samp[kv[0]]))[:70]:
        print(f"{key[0] + ':' + str(key[1]):34s} {v:10d} {100 * v / max(tot_e, 1):6.2f} {samp[key]:8d} "
              f"{100 * samp[key] / max(tot_s, 1):6.2f} {longsb[key]:8d}")


if __name__ == "__main__":
    main()


def by_function(rep, so, ksub, src, which=0, unit=1):
    """Aggregate executed warp-instructions per enclosing __device__ function of `src`."""
    dis = disasm_lines(so, ksub)
    hdr, body = ncu_sass(rep, ksub.replace("ILi", "<").split("<")[0], which)
    col = {h: i for i, h in enumerate(hdr)}
    fn_at, cur = {}, "?"
    for i, line in enumerate(open(src), 1):
        m = re.search(r"__device__\s+(?:__forceinline__|__noinline__|inline)?\s*[\w:<>\*& ]+?\s+(\w+)\s*\(", line)
        if m and not line.strip().startswith("//"):
            cur = m.group(1)
        g = re.search(r"__global__\s+void\s+(?:__launch_bounds__\([^)]*(?:\([^)]*\))?[^)]*\)\s*)?(\w+)\s*\(", line)
        if g and not line.strip().startswith("//"):
            cur = g.group(1) + " (kernel body)"
        fn_at[i] = cur
    base = os.path.basename(src)
    ex, samp = collections.Counter(), collections.Counter()
    for i in range(min(len(dis), len(body))):
        f, ln = dis[i][2]
        key = fn_at.get(ln, "?") if f == base else f
        ex[key] += int(body[i][col["Instructions Executed"]] or 0)
        samp[key] += int(body[i][col["# Samples"]] or 0)
    tot = sum(ex.values())
    print(f"{'function':32s} {'executed':>11s} {'%':>6s} {'per unit':>9s} {'samples%':>8s}")
    ts = sum(samp.values())
    for k, v in ex.most_common(40):
        print(f"{k:32s} {v:11d} {100 * v / tot:6.2f} {v / unit:9.1f} {100 * samp[k] / max(ts, 1):8.2f}")
