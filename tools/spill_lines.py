#!/usr/bin/env python
"""Where ptxas put local-memory spills: LDL/STL count per source line of one kernel (needs -lineinfo).
usage: python tools/spill_lines.py lib.so kernel_substring"""
import re, subprocess, sys, tempfile, os
from collections import Counter
lib, pat = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and pat in l)
end = next(i for i in range(start + 1, len(sass)) if sass[i].startswith(".text.") or sass[i].startswith(".section"))
cur, cnt = None, Counter()
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2)))
    if re.search(r"\b(LDL|STL)\b", l): cnt[(cur, "LDL" if "LDL" in l else "STL")] += 1
for k, v in sorted(cnt.items(), key=lambda x: (x[0][0][0], x[0][0][1])): print(k, v)
print("instructions:", sum(1 for l in sass[start:end] if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)))
