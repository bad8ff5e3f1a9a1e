#!/usr/bin/env python
"""Print the markdown results table of DESIGN.md §5 / README.md from a bench.py JSON line:
    python tools/bench_table.py profiles/r02_bench_default.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().splitlines()[-1])
r = d["roofline"]
e = d["e2e"]
c = d.get("cpu_baseline") or {}
print(f"| cfg2 (headline, {d['config']['workload'].split(':')[1].strip()[:60]}) | {d['value'] / 1e3:.1f} | {1e3 * d['ms_per_step'] / (d['config']['batches_per_step'] / d['config']['batches_per_engine_run']):.0f} us | "
      f"K1 {r['k1_ms']:.3f} / K2 {r['k2_ms']:.3f} | {r['frac']:.3f} | strict (tests) |")
for k, v in (d.get("configs") or {}).items():
    ps = v.get("parity_spot") or {}
    print(f"| {k} | {v['value'] / 1e3:.1f} | {1e3 * v['ms_per_engine_run']:.0f} us | K1 {v['k1_ms']:.3f} / K2 {v['k2_ms']:.3f} | {v.get('hbm_frac', 0):.3f} | {'ok' if ps.get('ok') else ps} ({(ps.get('mode') or '')[:7]}) |")
print()
print(f"e2e {e['value'] / 1e3:.1f} Gsamples/s, H2D {e.get('h2d_gbs_achieved', 0):.1f} GB/s = {e.get('pcie_frac', 0):.2f} of the measured pinned-copy rate; "
      f"CPU arm {c.get('value', 0) / 1e3:.2f} Gsamples/s on {c.get('cores')} threads ({c.get('kind')})")
print(f"roofline: achieved {r['achieved']:.0f} GB/s of {r['peak']:.0f} ({r['frac']:.3f}); traffic {r.get('traffic')}; issue {r.get('issue')}; tensor {r.get('tensor', {}).get('achieved_tops')}; k2 {r.get('k2')}")
