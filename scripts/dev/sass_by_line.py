#!/usr/bin/env python
"""Static SASS breakdown of one kernel by source file / line range (needs -lineinfo):
       scripts/dev/sass_by_line.py <kernel-name-substring> [lib.so]
Counts instructions per source file and per opcode class; for sample.cuh also per function (by line range)."""
import collections
import os
import re
import subprocess
import sys
import tempfile

pat = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "mahakala_b200", "libmahakala_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=tmp, capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(txt) if l.startswith(".text.") and pat in l]
if not start:
    sys.exit("kernel not found")
i0 = start[0]
i1 = next((i for i in range(i0 + 1, len(txt)) if txt[i].startswith(".text.") or txt[i].startswith(".section")), len(txt))
cur = None
byfile = collections.Counter()
byline = collections.Counter()
kinds = collections.defaultdict(collections.Counter)
for ln in txt[i0:i1]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
    if m and cur:
        op = m.group(2).split('.')[0]
        byfile[cur[0]] += 1
        byline[cur] += 1
        kinds[cur[0]][op] += 1
print(txt[i0])
print("total", sum(byfile.values()), dict(byfile))
for f in kinds:
    print(f"  {f}: {dict(kinds[f].most_common(16))}")
if "--lines" in sys.argv:
    for (f, l), v in sorted(byline.items()):
        print(f"{f}:{l}  {v}")
