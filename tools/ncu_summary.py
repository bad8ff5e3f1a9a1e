#!/usr/bin/env python
"""Summarise `ncu --set full` captures (gpurun_out/*.ncu-rep) into the tracked evidence under profiles/:

    python tools/ncu_summary.py --out profiles/r02_ncu_summary.txt --captures profiles/k1_captures.json \
        gpurun_out/r02_k1tc_cfg2.ncu-rep:cfg2:3 gpurun_out/r02_k2_cfg3.ncu-rep:cfg3 ...

Each argument is  <report>[:<workload>[:<fft_path 1|2|3, or k2 for a K2 capture>]].  For every kernel in a report the selected raw metrics are printed, plus
the ten hottest SASS instructions by stall samples (source page).  Reports of a K1 kernel given with workload and fft_path also
produce an entry of k1_captures.json: DRAM bytes and warp instructions per launch stamped with the sha of the kernel source, which
is what bench.py needs to put a measured `traffic` / `issue_frac` into its roofline (it refuses captures of other source versions).
"""
import argparse
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "smsp__inst_executed.sum", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]
SRC_OF = {1: "rtlsdr-airband_b200/csrc/k1_fft.cu", 2: "rtlsdr-airband_b200/csrc/k1_pruned.cu", 3: "rtlsdr-airband_b200/csrc/k1_tc.cu", "k2": "rtlsdr-airband_b200/csrc/k2_demod.cu"}


def sha_of(rel):
    return hashlib.sha256(open(os.path.join(ROOT, rel), "rb").read()).hexdigest()[:16]


def ncu(rep, page):
    r = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True)
    return r.stdout


def raw_rows(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def hot_instructions(rep, top=10):
    rows = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            n = int(r[isamp])
        except ValueError:
            continue
        stalls = {h.replace("stall_", ""): int(r[j]) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[j] not in ("0", "")}
        data.append((n, int(r[iex] or 0), r[isrc].strip(), stalls))
    total = sum(d[0] for d in data)
    out = [f"    stall samples: {total}; hottest instructions:"]
    for n, ex, src, st in sorted(data, key=lambda d: -d[0])[:top]:
        main = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        out.append(f"      {n:6d} ({100.0 * n / max(total, 1):4.1f} %)  executed {ex:9d}  {src[:70]:70s} [{main}]")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--out", required=True)
    ap.add_argument("--captures", default=None)
    ap.add_argument("--title", default="ncu --set full --clock-control none --import-source on")
    args = ap.parse_args()
    lines = [f"# {args.title}", "# (times under ncu are serialised and cold-cache: compare shares and per-launch counters, not absolute durations)", ""]
    captures = []
    if args.captures and os.path.exists(args.captures):
        captures = json.load(open(args.captures))
    for spec in args.reports:
        parts = spec.split(":")
        rep, workload = parts[0], (parts[1] if len(parts) > 1 else None)
        path = (parts[2] if parts[2] == "k2" else int(parts[2])) if len(parts) > 2 else None
        if not os.path.exists(rep):
            lines.append(f"## {rep}: missing")
            continue
        rows, units = raw_rows(rep)
        for row in rows:
            name = row.get("Kernel Name", "?")
            lines.append(f"## {os.path.basename(rep)}  workload={workload}  kernel={name[:100]}")
            for k in KEYS:
                if k in row and row[k] != "":
                    lines.append(f"    {k:95s} {row[k]:>16s} {units.get(k, '')}")
            try:
                dur_us = float(row["gpu__time_duration.sum"])
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                dram = sum(float(row[m]) * scale.get(units.get(m, ""), 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                inst = float(row.get("smsp__inst_executed.sum", "nan"))
                lines.append(f"    -> DRAM bytes per launch {dram:.0f}; warp instructions per launch {inst:.0f}")
                if path and workload:
                    captures = [c for c in captures if not (c.get("workload") == workload and c.get("fft_path") == path)]
                    captures.append({"workload": workload, "fft_path": path, "kernel": name[:80], "source": SRC_OF[path], "source_sha": sha_of(SRC_OF[path]),
                                     "dram_bytes_per_launch": dram, "warp_instructions_per_launch": inst, "duration_us_under_ncu": dur_us,
                                     "file": os.path.basename(args.out) + " <- " + os.path.basename(rep)})
            except Exception as ex:  # noqa: BLE001
                lines.append(f"    (could not derive per-launch numbers: {ex})")
        try:
            lines.extend(hot_instructions(rep))
        except Exception as ex:  # noqa: BLE001
            lines.append(f"    (source page unavailable: {ex})")
        lines.append("")
    open(args.out, "w").write("\n".join(lines) + "\n")
    if args.captures:
        json.dump(captures, open(args.captures, "w"), indent=1)
    print(f"wrote {args.out} ({len(lines)} lines)" + (f" and {args.captures} ({len(captures)} captures)" if args.captures else ""))


if __name__ == "__main__":
    main()
