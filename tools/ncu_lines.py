#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep (needs -lineinfo).

    python tools/ncu_lines.py gpurun_out/x/prof.ncu-rep <kernel regex> [top N]
Inlined helpers are attributed to the line of the helper (ncu's view); the listing is sorted by warp instructions executed.
"""
import csv
import subprocess
import sys


def main(rep, rx, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    out, seen, cur, hdr = [], set(), None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            if cur in seen:
                break            # second captured launch of the kernel starts here
            seen.add(cur)
            continue
        if r[0] == "Line No":
            hdr = r
            c_inst, c_samp, c_thr = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or len(r) <= c_thr or not r[0].strip().isdigit():
            continue
        try:
            out.append((int(r[c_inst]), int(r[c_samp]), int(r[c_thr]), cur + ":" + r[0], r[1].strip()))
        except ValueError:
            continue
    tot = sum(o[0] for o in out) or 1
    tots = sum(o[1] for o in out) or 1
    out.sort(reverse=True)
    print(f"total warp inst {tot}, samples {tots}")
    print("| line | warp inst | share | stall samples | share | avg active | source |")
    print("|---|---|---|---|---|---|---|")
    for inst, samp, thr, line, src in out[:top]:
        print(f"| {line} | {inst} | {100*inst/tot:.1f}% | {samp} | {100*samp/tots:.1f}% | {thr/max(inst,1):.1f} | `{src[:110]}` |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
