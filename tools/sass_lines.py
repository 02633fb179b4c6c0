#!/usr/bin/env python
"""Static SASS instruction count per CUDA source line of one kernel.   usage: python tools/sass_lines.py <cubin> <kernel name substring> [rows]
(cubins: `cuobjdump -xelf all laminr_b200/liblaminr_b200.so`; needs -lineinfo, which csrc/build.py passes).  The latency-bound one-shot kernels
run through their code once per launch with a cold instruction cache: code size is part of their time."""
import collections, re, subprocess, sys
out = subprocess.run(["nvdisasm", "-g", sys.argv[1]], capture_output=True, text=True).stdout
cnt, cur, intext, total = collections.Counter(), None, False, 0
for line in out.splitlines():
    if line.lstrip().startswith(".section"):
        intext = ".text." in line and sys.argv[2] in line
        continue
    if not intext:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+[A-Z@]", line):
        cnt[cur] += 1
        total += 1
print("SASS instructions:", total)
for k, v in cnt.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 25):
    print(f"{v:6d}  {k[0]}:{k[1]}")
