#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.md  ["title"]
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep profiles/r1_prof.md      ["title"]

`launches`: per-kernel count / average duration / share of the summed GPU time of a
    `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list.
`report`:  the headline counters of every kernel in a `ncu --set full` report (read with `ncu -i ... --page raw --csv`).
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]
KEEP_SUBSTR = ["pipe_tensor", "sm__inst_executed_pipe_xu_realtime"]


def launches(src, dst, title):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(H):
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nSource: `{src}` (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised "
                "-- read the SHARES, not the absolutes).\n\n| kernel | launches | avg us | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k[:110]}` | {a[0]} | {a[1] / a[0]:.2f} | {a[1] / tot:.3f} |\n")
        f.write(f"\nTotal GPU time in the list: {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches.\n")


def report(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    ki = H.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nSource: `{src}` (ncu --set full --clock-control none --import-source on), read with `ncu -i ... --page raw --csv`.\n")
        for r in rows[2:]:
            f.write(f"\n## `{r[ki][:120]}`\n\n| metric | unit | value |\n|---|---|---:|\n")
            for h, u, v in zip(H, U, r):
                if h in KEEP or any(s in h and h.endswith(("pct_of_peak_sustained_elapsed", "pct_of_peak_sustained_active")) and ".avg." in h for s in KEEP_SUBSTR):
                    f.write(f"| {h} | {u} | {v} |\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else dst
    {"launches": launches, "report": report}[mode](src, dst, title)
