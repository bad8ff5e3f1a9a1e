"""Kernel durations inside the steady-state pipeline (K1 of run i+1 overlapping K2 of run i) vs each kernel alone.
Run on the GPU box:  python tools/overlap_probe.py [workload]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'rtlsdr-airband_b200', 'py'), ROOT]
import numpy as np, torch
from airband_b200 import lib
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg, _ = bench.make_workload(wl)
nb = 4
raws = bench.synth_streams(cfg, nb)
eng = lib.Engine(cfg, max_batches_per_run=nb)
for d in range(len(cfg.devices)):
    eng.resident_load(d, raws[d])
for _ in range(5):
    eng.run_resident(nb)
eng.sync()
alone = []
for _ in range(5):
    eng.run_resident(nb)
    alone.append(eng.last_run_times())
a = np.median(np.array(alone), axis=0)
print("alone     : K1 %.3f  K2 %.3f  tail %.3f  run %.3f ms" % tuple(a))
piped = []
for _ in range(8):
    t = time.perf_counter()
    for _ in range(20):
        eng.run_resident(nb)
    piped.append(eng.last_run_times())
    eng.sync()
    dt = (time.perf_counter() - t) / 20
p = np.median(np.array(piped), axis=0)
print("pipelined : K1 %.3f  K2 %.3f  tail %.3f  run %.3f ms   (period %.3f ms wall)" % (*p, dt * 1e3))
eng.sync()
t = time.perf_counter()
for _ in range(20):
    eng.run_resident(nb)
t_enq = (time.perf_counter() - t) / 20
tl = eng.timeline(8)
print("host enqueue time per run (20 runs, pipeline full): %.3f ms" % (t_enq * 1e3))
if os.environ.get("PROBE_VERBOSE"):
    print("timeline of the last 8 runs (ms): K1 start, K1 end, K2 start, K2 end, run end")
    for r in range(8):
        print("  run %d: " % r + "  ".join("%7.3f" % v for v in tl[r]))
print("overlapped (run 3 of 8): K1 %.3f  K2 %.3f  period %.3f ms" % (tl[3][1] - tl[3][0], tl[3][3] - tl[3][2], tl[4][0] - tl[3][0]))
