#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small, tracked files under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_xxx_launches.md
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r1_xxx_full.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput",
        "dram__cycles_active", "sm__pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma", "l1tex__t_bytes",
        "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__shared_mem_per_block",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform", "smsp__warp_issue_stalled",
        "smsp__average_warps_issue_stalled", "smsp__issue_active.avg", "sm__pipe_fma_cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__waves_per_multiprocessor", "smsp__warps_eligible.avg")


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    return name.strip()


def launches(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), r["Grid Size"], float(r["Metric Value"])))
    agg = collections.OrderedDict()
    for k, g, ns in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("Per-launch times are cold-cache and serialised (ncu); compare SHARES, not absolutes.\n\n")
        f.write(f"{len(rows)} launches, total {total/1e6:.3f} ms\n\n| kernel | launches | total us | share | avg us |\n|---|--:|--:|--:|--:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns/1e3:.1f} | {100*ns/total:.1f}% | {ns/1e3/n:.1f} |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows:
            d = dict(zip(hdr, r))
            f.write(f"## {short(d['Kernel Name'])}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n| metric | value | unit |\n|---|--:|---|\n")
            for h, u, v in zip(hdr, units, r):
                if any(h.startswith(k) for k in KEYS) and not re.search(r"\.(min|max|sum)\.pct|per_second|_allocated|_driver|_static", h):
                    f.write(f"| {h} | {v} | {u} |\n")
            f.write("\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
