"""Static evidence, no GPU needed: per kernel of laminr_b200/liblaminr_b200.so, the resource usage (`cuobjdump -res-usage`) and the counts of the SASS
mnemonics that identify the Blackwell paths (`cuobjdump -sass`; /opt/skills/guides/B200_PROFILING.md, "What proves a Blackwell-native kernel"):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st, UTCBAR = tcgen05.commit, UBLKCP / UTMALDG = bulk (TMA) copies, SYNCS = mbarrier operations,
LDGSTS = cp.async, MUFU = special-function unit, HMMA = legacy mma.sync (must be 0).

    python tools/sass_summary.py [kernel-name substrings ...]        -> markdown on stdout
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "laminr_b200", "liblaminr_b200.so")
KEYS = ["UTC*MMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "MUFU", "HMMA", "LDL/STL", "total"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    want = sys.argv[1:]
    res = {}
    cur = None
    for line in subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            res[cur] = dict(re.findall(r"([A-Z]+)(?:\[0\])?:(\d+)", line))
            cur = None
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if not (cur and m):
            continue
        op = m.group(1)
        c = counts[cur]
        c["total"] += 1
        if re.match(r"UTC[A-Z]*MMA", op):
            c["UTC*MMA"] += 1
        elif op in ("LDL", "STL"):
            c["LDL/STL"] += 1
        elif op in KEYS:
            c[op] += 1
    names = sorted(counts)
    dm = demangle(names)
    print("| kernel | regs | stack B | " + " | ".join(KEYS) + " |")
    print("|---|---:|---:|" + "---:|" * len(KEYS))
    for n in names:
        pretty = re.sub(r"\(.*\)$", "", dm.get(n, n)).replace("void ", "")
        if want and not any(w in pretty for w in want):
            continue
        r = res.get(n, {})
        print(f"| `{pretty}` | {r.get('REG', '?')} | {r.get('STACK', '?')} | " + " | ".join(str(counts[n][k]) for k in KEYS) + " |")


if __name__ == "__main__":
    main()
