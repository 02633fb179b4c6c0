#!/usr/bin/env python
"""Source-level hot spots of one kernel of an `ncu --set full --import-source on` report.

    python tools/ncu_source_hot.py report.ncu-rep <launch index in the report> [rows]

Reads `ncu -i ... --page source --csv --print-source cuda,sass`, keeps the CUDA-source rows (a row per source line and file; inlined helpers are
attributed to their own lines) and prints them by stall samples: share of samples, share of executed warp instructions, the dominant stall reasons."""
import csv
import subprocess
import sys


def main():
    rep, idx = sys.argv[1], int(sys.argv[2])
    rows_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows, header, fname, func = [], None, "", ""
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            func = r[1]
        elif r[0] == "Line No":
            header = r
        elif header is not None and len(r) == len(header) and r[2] == "-":  # a CUDA-source row (no SASS address)
            d = dict(zip(header[4:], r[4:]))
            rows.append((fname, r[0], r[1].strip(), d))
    num = lambda d, k: float(d.get(k, 0) or 0)
    tot_s = sum(num(d, "# Samples") for *_, d in rows) or 1
    tot_i = sum(num(d, "Instructions Executed") for *_, d in rows) or 1
    print(f"kernel: {func}\nCUDA-source rows: {len(rows)}, stall samples {tot_s:.0f}, executed warp instructions (rows summed) {tot_i:.0f}\n")
    print("| samples | instr | line | top stall reasons | source |\n|---:|---:|---|---|---|")
    stall_keys = [k for k in header[4:] if k.startswith("stall_") and "Not Issued" not in k]
    for fname, ln, src, d in sorted(rows, key=lambda x: -num(x[3], "# Samples"))[:rows_n]:
        reasons = sorted(((num(d, k), k[6:]) for k in stall_keys), reverse=True)[:3]
        rs = " ".join(f"{k}:{v / max(num(d, '# Samples'), 1):.0%}" for v, k in reasons if v > 0)
        print(f"| {num(d, '# Samples') / tot_s:.1%} | {num(d, 'Instructions Executed') / tot_i:.1%} | `{fname}:{ln}` | {rs} | `{src[:110]}` |")


if __name__ == "__main__":
    main()
