"""Multi-GPU sharding of the demodulation path: devices[] is block-partitioned over ranks exactly like the reference
partitions it over demod threads (init_demod(device_start, device_end), reference src/rtl_airband.cpp:1070-1086).
Devices are independent, so there is no data-path collective; the only cross-device flow is a mixer whose inputs live
on several ranks (reference src/mixer.cpp:189-214): each rank sums its local inputs on its GPU and ONE small
all-reduce adds the partial sums (SUM) and ORs the has_signal flags (MAX)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

from .config import Config

MixerInput = Tuple[int, int, float, float]  # (dev, chan, ampfactor, balance)


def device_range(n_devices: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of devices owned by `rank`."""
    return (n_devices * rank) // world, (n_devices * (rank + 1)) // world


def shard_config(cfg: Config, rank: int, world: int) -> Config:
    s, e = device_range(len(cfg.devices), rank, world)
    return Config(fft_size=cfg.fft_size, wave_rate=cfg.wave_rate, fm_demod=cfg.fm_demod, devices=cfg.devices[s:e])


def shard_mixers(mixers: Sequence[Sequence[MixerInput]], n_devices: int, rank: int, world: int) -> List[List[MixerInput]]:
    """Per mixer, the inputs whose device lives on this rank, with device indices made engine-local.  Every rank keeps
    every mixer (possibly with no local input) so that the all-reduce buffers line up."""
    s, e = device_range(n_devices, rank, world)
    return [[(d - s, c, a, b) for (d, c, a, b) in m if s <= d < e] for m in mixers]


def allreduce_mixers(sums, flags, group=None):
    """In-place cross-rank reduction of the partial mixer sums (any torch tensor: CUDA with NCCL, CPU with gloo)."""
    import torch.distributed as dist
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    return sums, flags


class _DevPtr:
    """Minimal __cuda_array_interface__ carrier so torch can wrap engine-owned device memory without a copy."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3}


def engine_mixer_tensors(engine, n_mixers: int):
    """torch views of the engine's device-side mixer buffers: sums float32[nbmax, n_mixers, 2, B], flags int32[nbmax, n_mixers]."""
    import torch
    sums_ptr, flags_ptr = engine.mixer_device_buffers()
    sums = torch.as_tensor(_DevPtr(sums_ptr, (engine.nbmax, n_mixers, 2, engine.B), "<f4"), device="cuda")
    flags = torch.as_tensor(_DevPtr(flags_ptr, (engine.nbmax, n_mixers), "<i4"), device="cuda")
    return sums, flags
