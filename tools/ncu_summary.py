#!/usr/bin/env python
"""Summarise ncu output into small text files that can be committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.md
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep            > profiles/rNN_kernels.md
    python tools/ncu_summary.py sass     gpurun_out/prof.ncu-rep <kernel regex>   (opcode mix of one kernel)
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_static", "smem_static"),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_thr%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_thr%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_thr%"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("lts__t_sectors.sum", "l2_sectors"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


def short(name):
    name = name.replace("sgi::<unnamed>::", "").replace("void ", "")
    return name.split("(")[0]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    i_name, i_metric, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[1:]:
        if r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        u = r[i_unit]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        k = short(r[i_name])
        tot[k] = tot.get(k, 0.0) + us
        cnt[k] += 1
    total = sum(tot.values())
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {cnt[k]} | {v:.1f} | {v / cnt[k]:.2f} | {100 * v / total:.1f}% |")
    print(f"\ntotal {total:.1f} us over {sum(cnt.values())} launches (ncu per-launch times are cold-cache and serialised: compare shares, not absolutes)")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(k, n) for k, n in KEYS if k in idx]
    print("| kernel | " + " | ".join(n for _, n in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        cells = []
        for k, _ in cols:
            v, u = r[idx[k]], units[idx[k]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        print(f"| {short(r[idx['Kernel Name']])} | " + " | ".join(cells) + " |")


def sass(path, regex):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + regex, "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r]
    for n, hi in enumerate(starts[:1]):
        hdr = rows[hi]
        i_src, i_ex, i_th = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        tot, tt = collections.Counter(), 0
        end = starts[n + 1] if n + 1 < len(starts) else len(rows)
        for r in rows[hi + 1:end]:
            if len(r) <= i_th:
                continue
            try:
                c = int(r[i_ex])
            except ValueError:
                continue
            toks = r[i_src].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            tot[op.split(".")[0]] += c
            try:
                tt += int(r[i_th])
            except ValueError:
                pass
        s = sum(tot.values())
        print(f"kernel regex {regex}: {s} warp instructions, {tt} thread instructions ({tt / max(s, 1):.1f} active threads/warp-inst)")
        print("| opcode | warp inst | share |")
        print("|---|---|---|")
        for k, v in tot.most_common(20):
            print(f"| {k} | {v} | {100 * v / s:.1f}% |")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "full":
        full(sys.argv[2])
    else:
        sass(sys.argv[2], sys.argv[3])
