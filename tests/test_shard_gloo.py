"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: block partition of devices[], engine-local mixer input
lists, and the single all-reduce that completes mixers spanning ranks.  The per-rank audio comes from the oracle (a
stand-in for the per-GPU engines, which cannot run here); the result must equal the un-sharded mix."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_py as op
from airband_b200 import shard
from airband_b200 import workloads as wl


def test_device_ranges_partition():
    for n in (1, 3, 4, 64, 4096, 7):
        for w in (1, 2, 3, 8):
            r = [shard.device_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [e - s for s, e in r]
            assert max(sizes) - min(sizes) <= 1


def _partials(cfg, outs, mixers_local, nb):
    B = cfg.wave_batch
    sums = np.zeros((nb, len(mixers_local), 2, B), np.float32)
    flags = np.zeros((nb, len(mixers_local)), np.int32)
    for m, inputs in enumerate(mixers_local):
        for b in range(nb):
            for (d, c, amp, bal) in inputs:
                wo, _, ax = outs[d]
                if ax[b, c] == ord(' '):
                    continue
                flags[b, m] = 1
                ampl, ampr = np.float32(min(1.0, 1.0 - bal)), np.float32(min(1.0, 1.0 + bal))
                x = wo[c, b * B:(b + 1) * B]
                sums[b, m, 0] += x * (np.float32(amp) * ampl)
                sums[b, m, 1] += x * (np.float32(amp) * ampr)
    return sums, flags


def _worker(rank, world, port, nb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = wl.cfg4()
    mixers = wl.mixers_cfg4(cfg)
    mixers = [m[1] for m in mixers]
    s, e = shard.device_range(len(cfg.devices), rank, world)
    sub = shard.shard_config(cfg, rank, world)
    raws = [wl.synth_iq(cfg, d, wl.samples_for_batches(cfg, d, nb), key_on_s=0.2, key_off_s=0.1) for d in range(s, e)]
    outs, _ = op.run_oracle(sub, raws)
    local = shard.shard_mixers(mixers, len(cfg.devices), rank, world)
    sums, flags = _partials(sub, outs, local, nb)
    ts, tf = torch.from_numpy(sums), torch.from_numpy(flags)
    shard.allreduce_mixers(ts, tf)
    if rank == 0:
        q.put((ts.numpy().copy(), tf.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_mixer_allreduce_world2_equals_single_process():
    nb, world = 2, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    got_s, got_f = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = wl.cfg4()
    mixers = [m[1] for m in wl.mixers_cfg4(cfg)]
    raws = [wl.synth_iq(cfg, d, wl.samples_for_batches(cfg, d, nb), key_on_s=0.2, key_off_s=0.1) for d in range(len(cfg.devices))]
    outs, _ = op.run_oracle(cfg, raws)
    ref_s, ref_f = _partials(cfg, outs, mixers, nb)
    assert np.array_equal(got_f, ref_f) and ref_f.any()
    assert np.allclose(got_s, ref_s, atol=1e-6)  # summation order across ranks differs at the 1e-7 level
