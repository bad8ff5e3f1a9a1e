#!/usr/bin/env python
"""bench.py — throughput of the multichannel demodulation hot path (BASELINE.json metric:
"IQ Msamples/s through FFT+demod at 1/2/4/8 B200; % HBM roofline; vs CPU ref").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg1|cfg3|cfg5] [--impl reference]

A step = one pass of the hot path (K1 convert+window+FFT+bins, K2 demodulation) over one batch of synthetic input:
`batches_per_step` WAVE_BATCHes (default 4 x 125 ms) of every device of the workload.  Default workload at every
N is BASELINE.json configs[1] per GPU: 64 devices x 2.56 Msps U8, fft_size 2048, 8 AM channels each ("cfg2");
devices shard by GPU with no data-path collective, so N GPUs run N x 64 devices (scaling = "weak").

  value  device-timed (CUDA events on the engine's stream, max over ranks): IQ samples consumed / s, inputs resident
         in HBM (the resident stream, 168 MB per step, is larger than the 126 MB L2, so every step re-reads HBM).
  e2e    same metric through the public C ABI with HOST buffers: abg_push (H2D from pinned memory) + abg_run +
         abg_fetch_batches (results written by the GPU into pinned host slots, then copied to the caller's arrays) inside
         the timed region, software-pipelined by one step like any streaming caller (wall clock around the loop).
  clocks nvidia-smi SM clock / throttle reasons sampled every 20 ms during the timed region (the sampler's first row is
         awaited before timing starts; a region shorter than 5 samples is extended by identical untimed steps).
  roofline     K1 (the dominant kernel): algorithmic bytes per launch / CUDA-event duration vs the measured HBM
               peak; the FP32 figures next to it are the binding ones for this path.
  cpu_baseline the CPU oracle (reference leaf classes + restated loop, the reference's own -O3 -ffast-math flags)
               on this box's host cores over a bounded sample of the same workload.
`--impl reference` times that CPU path alone (rank 0 only under torchrun) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "rtlsdr-airband_b200", "py"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "iq_msamples_per_s_fft_demod"
UNIT = "Msamples/s"


def make_workload(name: str):
    from airband_b200 import config as cm
    from airband_b200 import workloads as wl
    if name == "cfg2":
        cfg = wl.cfg2(n_devices=64, n_channels=8)
        desc = "cfg2: 64 synthetic devices x 2.56 Msps U8, fft_size 2048, 8 AM channels each (BASELINE.json configs[1])"
    elif name == "cfg1":
        cfg = wl.cfg1()
        desc = "cfg1: 1 device, 2.56 Msps U8, fft_size 512, 1 AM channel (config/basic_multichannel.conf shape)"
    elif name == "cfg3":
        cfg = wl.cfg3(n_devices=8, n_channels=32, sfmt=cm.SFMT_S16)
        desc = "cfg3: 8 devices x 10 Msps S16, NFM build (WAVE_RATE 16000), fft_size 4096, 32 NFM channels with CTCSS+notch"
    elif name == "cfg5":
        cfg = wl.cfg5(n_devices=512, n_channels=8)
        desc = "cfg5 (one GPU's share): 512 devices x 2.56 Msps U8, fft_size 512, 8 AM channels each"
    else:
        raise SystemExit(f"unknown workload {name}")
    return cfg, desc


def synth_streams(cfg, n_batches: int, n_unique: int = 4):
    """Synthetic raw streams: n_unique distinct seeded streams tiled over the devices."""
    from airband_b200 import workloads as wl
    uniq = []
    for u in range(min(n_unique, len(cfg.devices))):
        n = wl.samples_for_batches(cfg, u, n_batches)
        uniq.append(wl.synth_iq(cfg, u, n, key_on_s=0.30, key_off_s=0.12))
    return [uniq[d % len(uniq)] for d in range(len(cfg.devices))]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout_s: float = 10.0):
        """nvidia-smi needs a few hundred ms to produce its first row: the timed region must not start before that."""
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout_s:
            time.sleep(0.01)

    def count_between(self, t0: float, t1: float) -> int:
        return sum(1 for (t, _) in self.rows if t0 <= t <= t1)

    def stop(self, windows=None):
        """Summary over the samples whose timestamp lies in one of `windows` [(t0, t1), ...] (None: all samples)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for (ts, r) in self.rows:
            if windows is not None and not any(a <= ts <= b for (a, b) in windows):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm),
                "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def cpu_run(cfg, desc, steps: int, warmup: int, budget_s: float = 20.0):
    """Time the CPU path (oracle) over bounded samples of the workload. Returns (Msps, info dict, ms_per_step)."""
    import oracle_py as op
    from airband_b200 import workloads as wl
    variant = "ref_fast" if op.available("ref_fast") else "restated_fast"
    kind = "reference" if variant == "ref_fast" else "port"
    cores = os.cpu_count() or 1
    D = min(len(cfg.devices), max(1, cores))
    sub = type(cfg)(fft_size=cfg.fft_size, wave_rate=cfg.wave_rate, fm_demod=cfg.fm_demod, devices=cfg.devices[:D])
    nthreads = min(cores, D)
    hop = sub.hop(0)
    B = sub.wave_batch
    # calibrate: one batch per device
    o = op.Oracle(sub, variant)
    o.set_discard(True)
    raws = synth_streams(sub, 1 + 1)
    for d in range(D):
        o.push(d, raws[d])
    t0 = time.perf_counter()
    nb0 = o.run(1, nthreads)
    dt0 = max(time.perf_counter() - t0, 1e-4)
    o.close()
    per_chunk = dt0 / max(nb0, 1) * D * 4            # seconds for "all sample devices advance 4 batches"
    total_steps = steps + warmup
    nb_chunk = 4
    reps = max(1, min(64, int(budget_s / max(per_chunk * total_steps, 1e-9))))
    o = op.Oracle(sub, variant)
    o.set_discard(True)
    raws = synth_streams(sub, nb_chunk)
    chunk_items = [nb_chunk * B * sub.hop(d) * 2 for d in range(D)]
    prime_items = [(100 * sub.hop(d) + sub.fft_size) * 2 for d in range(D)]
    for d in range(D):  # priming part once; afterwards the same 4 batches of samples are replayed (pushes are untimed)
        o.push(d, raws[d][:prime_items[d]])
    done, dt = 0, 0.0
    for step in range(total_steps):
        for _ in range(reps):
            for d in range(D):
                o.push(d, raws[d][prime_items[d]:prime_items[d] + chunk_items[d]])
            t0 = time.perf_counter()
            n = o.run(nb_chunk, nthreads)
            t1 = time.perf_counter()
            if step >= warmup:
                done += n
                dt += t1 - t0
    o.close()
    nb_step = nb_chunk * reps
    samples = done * B * hop
    msps = samples / dt / 1e6
    info = {"value": msps, "unit": UNIT, "cores": nthreads, "kind": kind,
            "sample": f"{D} of {len(cfg.devices)} devices x {nb_step} batches/step x {steps} steps ({samples / 1e6:.1f} Msamples, {dt:.1f} s); "
                      f"oracle variant {variant} (-O3 -ffast-math, x86-64-v3; FFTW unavailable offline -> own FP32 FFT), one thread per device"}
    return msps, info, dt / steps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--batches-per-step", type=int, default=4)
    ap.add_argument("--fft-mode", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg, desc = make_workload(args.workload)

    # ------------------------------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        msps, info, ms = cpu_run(cfg, desc, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": msps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "note": "CPU reference arm: bounded sample of the workload on host cores"},
                "cpu_baseline": info, "e2e": {"value": msps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from airband_b200 import lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nb = args.batches_per_step
    eng = lib.Engine(cfg, cuda_device=local_rank, max_batches_per_run=nb, input_capacity_batches=nb + 1, fft_mode=args.fft_mode)
    stream = torch.cuda.Stream(device=local_rank)
    eng.set_stream(stream.cuda_stream)
    B = eng.B
    D = len(cfg.devices)
    hop = [cfg.hop(d) for d in range(D)]
    bpc = [2 * cfg.devices[d].bytes_per_sample for d in range(D)]
    raws = synth_streams(cfg, nb)
    for d in range(D):
        eng.resident_load(d, raws[d])
    samples_per_step = sum(nb * B * hop[d] for d in range(D))
    resident_bytes = sum(eng.resident_bytes_needed(d) for d in range(D))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-timed, inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for _ in range(max(args.warmup, 1)):
        eng.run_resident(nb)
    barrier()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_ms, k2_ms = [], []
    barrier()
    t_wall0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        eng.run_resident(nb)
    eng.join()  # main stream waits for the K2 stream: ev1 covers every kernel of every step
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    clocks = None
    if rank == 0:
        windows = [(t_wall0, t_wall1)]
        note = None
        if sampler.count_between(t_wall0, t_wall1) < 5:
            # the timed region is shorter than a handful of 20 ms sampling intervals (small --steps): keep the GPU under the
            # SAME load (identical untimed steps) until enough samples exist, and say so
            t_x0 = time.time()
            while sampler.proc and sampler.count_between(t_x0, time.time()) < 8 and time.time() - t_x0 < 3.0:
                for _ in range(50):
                    eng.run_resident(nb)
                eng.sync()
            windows.append((t_x0, time.time()))
            note = "timed region shorter than 5 sampling intervals: clocks also sampled over identical untimed steps run right after it"
        clocks = sampler.stop(windows)
        if note:
            clocks["note"] = note
    # per-kernel durations from the engine's own events (same stream), a few extra steps outside the timed loop
    for _ in range(5):
        eng.run_resident(nb)
        t = eng.last_run_times()
        k1_ms.append(t[0]); k2_ms.append(t[1])
    t_all = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t_all.item())
    value = world * samples_per_step * args.steps / (elapsed_ms * 1e-3) / 1e6

    # ---- e2e: host buffers through the public API (push H2D + run + fetch D2H) ----
    e2e = None
    if not args.no_e2e:
        # software-pipelined like any streaming user of the API: the input of step i+1 is pushed (H2D on the engine's
        # ingest stream) before the results of step i are fetched, so copies and kernels of neighbouring steps overlap
        eng2 = lib.Engine(cfg, cuda_device=local_rank, max_batches_per_run=nb, input_capacity_batches=2 * nb + 1, fft_mode=args.fft_mode)
        step_items = [nb * B * hop[d] * 2 for d in range(D)]          # array items per step per device
        prime_items = [(100 * hop[d] + cfg.fft_size) * 2 for d in range(D)]
        pinned = []
        for d in range(D):
            need = prime_items[d] + step_items[d]
            src = torch.from_numpy(np.ascontiguousarray(raws[d][:need])).pin_memory()
            pinned.append(src)
        wo = [np.empty((nb, len(cfg.devices[d].channels), B), np.float32) for d in range(D)]
        ax = [np.empty((nb, len(cfg.devices[d].channels)), np.uint8) for d in range(D)]
        item = [cfg.devices[d].bytes_per_sample for d in range(D)]

        def submit(first: bool):
            for d in range(D):
                base = pinned[d].data_ptr()
                if first:
                    eng2.push_ptr(d, base, (prime_items[d] + step_items[d]) * item[d])
                else:  # replay the same 4 batches of host samples (skipping the priming part)
                    eng2.push_ptr(d, base + prime_items[d] * item[d], step_items[d] * item[d])
            n = eng2.run(nb)
            assert n == D * nb, (n, D * nb)

        def collect():
            for d in range(D):
                assert eng2.fetch_many_into(d, nb, wo[d], ax[d]) == nb

        submit(True)
        for _ in range(max(args.warmup, 1)):
            submit(False)
            collect()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            submit(False)   # step i+1 in flight ...
            collect()       # ... while step i's results are fetched (every step's input and output cross PCIe in here)
        eng2.sync()
        barrier()
        dt = time.perf_counter() - t0
        collect()
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        G = sum(len(dv.channels) for dv in cfg.devices)
        e2e = {"value": world * samples_per_step * args.steps / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(sum(step_items[d] * item[d] for d in range(D))),
               "d2h_bytes_per_step": int(G * nb * B * 4 + nb * ((G + 31) // 32 * 32)),
               "timing": "wall clock around synchronised steps (includes host-side copies out of the pinned result slots)"}
        eng2.close()

    # ---- roofline of the dominant kernel (K1) ----
    peaks, peak_src = measured_peaks()
    N = cfg.fft_size
    frames_per_launch = D * nb * B
    b_alg = [hop[d] * bpc[d] + 4 * len(cfg.devices[d].channels) for d in range(D)]
    alg_bytes = float(sum(b_alg[d] * nb * B for d in range(D)))
    k1 = float(np.median(k1_ms)) * 1e-3
    k2 = float(np.median(k2_ms)) * 1e-3
    achieved = alg_bytes / k1 / 1e9
    flops_per_frame = 5 * N * math.log2(N) + 2 * N
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_dram_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    roofline = {"bound": "hbm", "kernel": ("k1_fft_kernel" if args.fft_mode == 1 else "k1_pruned_kernel") + " (convert+window+FFT+bin select)", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback 6650 GB/s",
                "traffic": traffic, "alg_bytes_per_launch": alg_bytes, "k1_ms": k1 * 1e3, "k2_ms": k2 * 1e3,
                "k1_share_of_step": k1 / max(k1 + k2, 1e-12),
                "fp32": {"achieved_tflops": frames_per_launch * flops_per_frame / k1 / 1e12, "peak_tflops_at_observed_clock": fp32_peak,
                         "frac": frames_per_launch * flops_per_frame / k1 / 1e12 / fp32_peak,
                         "note": "5*N*log2(N)+2N flop per frame; the path is FP32-issue bound for U8/S16 input (SURVEY.md §8d), so HBM frac is small by construction"}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": desc, "devices_per_gpu": D, "batches_per_step": nb, "wave_rate": cfg.wave_rate, "fft_mode": args.fft_mode,
                       "realtime_floor_msps_per_gpu": sum(dv.sample_rate for dv in cfg.devices) / 1e6,
                       "l2": f"resident input {resident_bytes / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)" if resident_bytes > 126e6
                       else f"resident input {resident_bytes / 1e6:.0f} MB per step fits L2: value is L2-warm"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            _, info, _ = cpu_run(cfg, desc, steps=3, warmup=1, budget_s=15.0)
            line["cpu_baseline"] = info
        except Exception as ex:  # the oracle is a checker; its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
