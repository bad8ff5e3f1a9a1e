#!/usr/bin/env python
"""Aggregate the warp-stall samples of an ncu report per CUDA source line (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py gpurun_out/r02_k2_cfg2.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if r and r[0] == "Line No")
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(j, h[6:]) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
cur = None
for r in rows:
    if not r or r[0] == "Line No" or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur = (int(r[0]), r[1].strip())
        continue
    if cur is None or r[2] in ("", "..."):
        continue
    try:
        n = int(r[isamp])
        ex = int(r[iex])
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0, 0, {}])
    a[0] += n
    a[1] += ex
    a[2] += 1
    for j, name in stall_cols:
        try:
            v = int(r[j])
        except ValueError:
            v = 0
        if v:
            a[3][name] = a[3].get(name, 0) + v
tot = sum(a[0] for a in agg.values())
totex = sum(a[1] for a in agg.values())
print(f"{rep}: {tot} samples, {totex} warp instructions executed")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{k} {v}" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
    print(f"{a[0]:7d} {100.0 * a[0] / max(tot, 1):5.1f}%  ex {100.0 * a[1] / max(totex, 1):5.1f}%  sass {a[2]:4d}  L{ln:<5d} {src[:90]:90s} [{st}]")
