"""The one collective of the path on real hardware (north_star "one NCCL collective for the shared mixer path"):
a mixer whose inputs live on two GPUs.  Each rank runs its own engine over its share of devices[], sums its local inputs on
its GPU (abg_mixer_device_buffers), one NCCL all-reduce adds the partial sums in place (SUM) and ORs the has_signal flags
(MAX); the result must equal the un-sharded engine's mixer output and the oracle-side sum (reference src/mixer.cpp:133-140,
189-214).  Needs two GPUs: skipped on a one-GPU box (the CPU/gloo version of the same logic is tests/test_shard_gloo.py)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
NB = 2


def _mixers(cfg):
    n_dev = len(cfg.devices)
    return [[(d, m, 1.0 + 0.25 * d, (-0.5 if (m == 1 and d == 0) else 0.0)) for d in range(n_dev)] for m in range(4)]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from airband_b200 import lib, shard
    from airband_b200 import workloads as wl
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = wl.cfg4()
    mixers = _mixers(cfg)
    s, e = shard.device_range(len(cfg.devices), rank, world)
    sub = shard.shard_config(cfg, rank, world)
    raws = [wl.synth_iq(cfg, d, wl.samples_for_batches(cfg, d, NB), key_on_s=0.2, key_off_s=0.1) for d in range(s, e)]
    eng = lib.Engine(sub, cuda_device=rank, max_batches_per_run=NB)
    eng.configure_mixers(shard.shard_mixers(mixers, len(cfg.devices), rank, world))
    for d, r in enumerate(raws):
        eng.push(d, r)
    assert eng.run(NB) == NB * (e - s)
    eng.sync()
    sums, flags = shard.engine_mixer_tensors(eng, len(mixers))
    assert sums.is_cuda and sums.device.index == rank
    shard.allreduce_mixers(sums, flags)          # NCCL, in place on the engine's own device buffers
    torch.cuda.synchronize()
    if rank == 0:
        q.put((sums.cpu().numpy().copy(), flags.cpu().numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_mixer_partial_sums_allreduced_by_nccl_equal_the_unsharded_engine_and_the_oracle():
    import torch.multiprocessing as mp
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got_s, got_f = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    import oracle_py as op
    import parity
    from airband_b200 import lib
    from airband_b200 import workloads as wl
    cfg = wl.cfg4()
    mixers = _mixers(cfg)
    raws = [wl.synth_iq(cfg, d, wl.samples_for_batches(cfg, d, NB), key_on_s=0.2, key_off_s=0.1) for d in range(len(cfg.devices))]
    # un-sharded engine on one GPU
    e = lib.Engine(cfg, cuda_device=0, max_batches_per_run=NB)
    e.configure_mixers(mixers)
    for d, r in enumerate(raws):
        e.push(d, r)
    assert e.run(NB) == NB * len(raws)
    B = cfg.wave_batch
    for m in range(4):
        for b in range(NB):
            left, right, sig = e.fetch_mixer(m)
            assert bool(got_f[b, m]) == sig
            assert np.allclose(got_s[b, m, 0], left, atol=1e-6) and np.allclose(got_s[b, m, 1], right, atol=1e-6)
    e.close()
    # oracle-side sum
    ores, _ = op.run_oracle(cfg, raws)
    ref = parity.mixer_reference(cfg, ores, mixers, NB)
    any_signal = False
    for m in range(4):
        for b in range(NB):
            left, right, sig = ref[m][b]
            any_signal |= sig
            assert bool(got_f[b, m]) == sig
            assert parity.gate(got_s[b, m, 0], left) <= parity.TOL and parity.gate(got_s[b, m, 1], right) <= parity.TOL
    assert any_signal
