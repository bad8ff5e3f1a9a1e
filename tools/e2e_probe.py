import sys, time, os
sys.path[:0]=['/root/repo/rtlsdr-airband_b200/py','/root/repo']
import numpy as np, torch
from airband_b200 import lib, workloads as wl
import bench
cfg,_=bench.make_workload("cfg2")
nb=4
raws=bench.synth_streams(cfg, nb)
D=len(cfg.devices)
# raw H2D bandwidth
x=torch.empty(164_000_000,dtype=torch.uint8).pin_memory(); y=torch.empty_like(x,device='cuda')
torch.cuda.synchronize()
for _ in range(3): y.copy_(x,non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): y.copy_(x,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print("H2D 164MB pinned: %.2f ms  %.1f GB/s"%(dt*1e3, 0.164/dt))
z=torch.empty(8_200_000,dtype=torch.uint8).pin_memory(); w=torch.empty_like(z,device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): z.copy_(w,non_blocking=True)
torch.cuda.synchronize(); print("D2H 8.2MB: %.3f ms"%((time.perf_counter()-t)/10*1e3))
eng=lib.Engine(cfg,max_batches_per_run=nb,input_capacity_batches=nb+1)
B=eng.B; hop=[cfg.hop(d) for d in range(D)]
step_items=[nb*B*hop[d]*2 for d in range(D)]; prime=[(100*hop[d]+cfg.fft_size)*2 for d in range(D)]
pinned=[torch.from_numpy(np.ascontiguousarray(raws[d][:prime[d]+step_items[d]])).pin_memory() for d in range(D)]
wo=[np.empty((8,B),np.float32) for d in range(D)]; ax=[np.empty(8,np.uint8) for d in range(D)]
def step(first, T):
    t0=time.perf_counter()
    for d in range(D):
        base=pinned[d].data_ptr()
        if first: eng.push_ptr(d,base,prime[d]+step_items[d])
        else: eng.push_ptr(d,base+prime[d],step_items[d])
    t1=time.perf_counter()
    n=eng.run(nb)
    t2=time.perf_counter()
    for d in range(D):
        for _ in range(nb): eng.fetch_into(d,wo[d],ax[d])
    t3=time.perf_counter()
    T[0]+=t1-t0; T[1]+=t2-t1; T[2]+=t3-t2
T=[0,0,0]; step(True,T)
for _ in range(3): step(False,T)
eng.sync(); T=[0,0,0]; t=time.perf_counter()
for _ in range(10): step(False,T)
eng.sync(); dt=(time.perf_counter()-t)/10
print("e2e step %.2f ms: push-enqueue %.2f run-enqueue %.2f fetch(wait+copy) %.2f"%(dt*1e3,T[0]*100,T[1]*100,T[2]*100))
