"""e2e timing probe: where does a pipelined push/run/fetch step spend its time?  (run on the GPU box)"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'rtlsdr-airband_b200', 'py'), ROOT]
import numpy as np, torch
from airband_b200 import lib
import bench
cfg, _ = bench.make_workload("cfg2")
nb = 4
raws = bench.synth_streams(cfg, nb)
D = len(cfg.devices)
x = torch.empty(164_000_000, dtype=torch.uint8).pin_memory(); y = torch.empty_like(x, device='cuda')
torch.cuda.synchronize()
for _ in range(3): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print("H2D 164MB pinned single copy: %.2f ms  %.1f GB/s" % (dt * 1e3, 0.164 / dt))
eng = lib.Engine(cfg, max_batches_per_run=nb, input_capacity_batches=2 * nb + 1)
B = eng.B; hop = [cfg.hop(d) for d in range(D)]
step_items = [nb * B * hop[d] * 2 for d in range(D)]; prime = [(100 * hop[d] + cfg.fft_size) * 2 for d in range(D)]
pinned = [torch.from_numpy(np.ascontiguousarray(raws[d][:prime[d] + step_items[d]])).pin_memory() for d in range(D)]
wo = [np.empty((8, B), np.float32) for d in range(D)]; ax = [np.empty(8, np.uint8) for d in range(D)]
T = [0.0] * 4
def submit(first):
    t0 = time.perf_counter()
    for d in range(D):
        base = pinned[d].data_ptr()
        if first: eng.push_ptr(d, base, prime[d] + step_items[d])
        else: eng.push_ptr(d, base + prime[d], step_items[d])
    t1 = time.perf_counter()
    eng.run(nb)
    t2 = time.perf_counter()
    T[0] += t1 - t0; T[1] += t2 - t1
def collect():
    t0 = time.perf_counter()
    eng.fetch_into(0, wo[0], ax[0])          # includes the wait for the run to finish
    t1 = time.perf_counter()
    for d in range(D):
        for k in range(nb):
            if d == 0 and k == 0: continue
            eng.fetch_into(d, wo[d], ax[d])
    t2 = time.perf_counter()
    T[2] += t1 - t0; T[3] += t2 - t1
submit(True)
for _ in range(3): submit(False); collect()
eng.sync()
for i in range(4): T[i] = 0.0
t = time.perf_counter()
for _ in range(10): submit(False); collect()
eng.sync(); dt = (time.perf_counter() - t) / 10
collect()
print("pipelined e2e step %.2f ms: push %.2f run %.2f first-fetch(wait) %.2f other-fetches(copy) %.2f" % (dt * 1e3, T[0] * 100, T[1] * 100, T[2] * 100, T[3] * 100))
# the same H2D volume alone through abg_push
eng3 = lib.Engine(cfg, max_batches_per_run=nb, input_capacity_batches=2 * nb + 1)
for d in range(D): eng3.push_ptr(d, pinned[d].data_ptr(), prime[d] + step_items[d])
eng3.run(nb); eng3.sync()
t = time.perf_counter()
for _ in range(5):
    for d in range(D): eng3.push_ptr(d, pinned[d].data_ptr() + prime[d], step_items[d])
    eng3.sync()
    eng3.run(nb)
    for d in range(D):
        for k in range(nb): eng3.fetch_into(d, wo[d], ax[d])
print("push-only phase measured inside a serial loop; total serial step %.2f ms" % ((time.perf_counter() - t) / 5 * 1e3))
t = time.perf_counter()
for d in range(D): eng3.push_ptr(d, pinned[d].data_ptr() + prime[d], step_items[d])
eng3.sync()
print("64 pushes + sync: %.2f ms" % ((time.perf_counter() - t) * 1e3))
