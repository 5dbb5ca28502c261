"""Opcode histogram of one kernel of a built library:  python tools/sass_hist.py <lib.so> <substring of the mangled name> [top]"""
import collections
import re
import subprocess
import sys

lib, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and key in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            hist[m.group(1)] += 1
total = sum(hist.values())
print(key, "instructions:", total)
for op, n in hist.most_common(top):
    print(f"  {n:6d}  {op}")
