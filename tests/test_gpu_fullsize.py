"""GPU parity at the FULL sizes of BASELINE.json's configs (the scaled-down shapes live in test_gpu_parity.py): every
device of the launch is compared bit for bit with its twin (devices that receive identical bytes must produce identical
audio and decisions wherever their tiles / warps ran), and the CPU oracle runs on one device per distinct synthetic
stream.  The throughput variant of configs[2] (squelch_snr_threshold = 0, what bench.py times) uses the relaxed comparison
SURVEY.md §7.3 prescribes: state-transition indices and audio compared separately (oracle/parity.py)."""
import numpy as np
import pytest

import bench
import oracle_py as op
import parity
from airband_b200 import config as cm
from airband_b200 import lib
from airband_b200 import workloads as wl

pytestmark = pytest.mark.gpu
NB = 4


def _twins_identical(res, n_unique):
    opened = 0
    for d in range(len(res)):
        w, _, a = res[d]
        w0, _, a0 = res[d % n_unique]
        assert np.array_equal(w.view(np.uint32), w0.view(np.uint32)) and np.array_equal(a, a0), f"device {d} differs from its twin {d % n_unique}"
        opened += int((a != ord(' ')).sum())
    return opened


def _oracle_on_distinct(cfg, raws, res, n_unique, relaxed=False):
    sub = cm.Config(fft_size=cfg.fft_size, wave_rate=cfg.wave_rate, fm_demod=cfg.fm_demod, devices=cfg.devices[:n_unique])
    ores, oorc = op.run_oracle(sub, raws[:n_unique])
    out = []
    for d in range(n_unique):
        r = (parity.relaxed if relaxed else parity.strict)(res[d], ores[d])
        assert r["ok"], (d, r)
        out.append(r)
    return out, oorc


@pytest.mark.parametrize("sfmt", [cm.SFMT_S16, cm.SFMT_F32], ids=["s16", "f32"])
def test_cfg3_full_size_parity_variant(sfmt):
    """BASELINE configs[2] at 8 devices x 32 NFM channels (CTCSS + notch + low-pass), manual -30 dBFS squelch: strict gate."""
    cfg = wl.cfg3(n_devices=8, n_channels=32, sfmt=sfmt, parity=True)
    raws = bench.synth_streams(cfg, NB, n_unique=4)
    res, eng = lib.demodulate_all(cfg, raws, max_batches_per_run=NB)
    assert res[0][0].shape == (32, NB * cfg.wave_batch)
    assert _twins_identical(res, 4) > 0
    per, oorc = _oracle_on_distinct(cfg, raws, res, 4)
    for d in range(4):
        for c in range(0, 32, 5):
            gs, os_ = eng.stats(d, c), oorc.stats(d, c)
            assert (gs.open_count, gs.ctcss_count, gs.no_ctcss_count, gs.active_counter) == (os_.open_count, os_.ctcss_count, os_.no_ctcss_count, os_.active_counter)
    eng.close()


def test_cfg3_full_size_throughput_variant_snr0():
    """The variant bench.py times: squelch_snr_threshold 0 (noaa.conf:24).  Transition indices and audio compared separately."""
    cfg, _ = bench.make_workload("cfg3")
    raws = bench.synth_streams(cfg, NB, n_unique=4)
    res, eng = lib.demodulate_all(cfg, raws, max_batches_per_run=NB)
    assert _twins_identical(res, 4) > 0
    per, _ = _oracle_on_distinct(cfg, raws, res, 4, relaxed=True)
    assert sum(p["audio_samples_compared"] for p in per) > 100000
    eng.close()


@pytest.mark.parametrize("lpw", [None, 8])
def test_cfg5_full_size_one_gpu_share(lpw, monkeypatch):
    """BASELINE configs[4], one GPU's share: 512 devices x 8 AM channels, fft 512 = 4096 channel chains.  The engine's
    channels-per-warp rule gives every chain its own warp (the lane-parallel tiles); the second case forces 8 channels per
    warp, the several-lanes-per-warp K2 builds that larger shares use, at the same full size."""
    if lpw is not None:
        monkeypatch.setenv("ABG_K2_LPW", str(lpw))
    cfg, _ = bench.make_workload("cfg5")
    assert sum(len(d.channels) for d in cfg.devices) == 4096
    raws = bench.synth_streams(cfg, NB, n_unique=4)
    res, eng = lib.demodulate_all(cfg, raws, max_batches_per_run=NB)
    assert len(res) == 512 and res[0][0].shape == (8, NB * cfg.wave_batch)
    assert _twins_identical(res, 4) > 0
    _oracle_on_distinct(cfg, raws, res, 4)
    eng.close()


def test_cfg4_baseline_shape_with_mixers():
    """BASELINE configs[3]: 4 devices x 4 AM channels into 4 mixers, every mixer spanning all devices; device audio and
    mixed sums vs the oracle-side sum (mixer.cpp:133-140,189-214)."""
    cfg, _ = bench.make_workload("cfg4")
    raws = bench.synth_streams(cfg, NB, n_unique=4)
    mixers = [m[1] for m in wl.mixers_cfg4(cfg)]
    r = bench.parity_spot(cfg, raws, NB, n_unique=4, mixers=mixers)
    assert r["ok"] and r["mixer_flags_equal"] and r["mixer_max_err"] <= parity.TOL, r
    assert r["opened"] > 0


@pytest.mark.parametrize("n,shift_bins,afc", [(512, 3, 2), (512, -3, 2), (1024, 2, 1), (2048, -4, 3)])
def test_afc_directions_and_sizes(n, shift_bins, afc):
    """AFC (reference src/rtl_airband.cpp:180-251) up and down, three FFT sizes: the bin must follow the carrier on the
    squelch-open edge exactly as in the oracle, batch by batch."""
    sr, w, cf = 2560000, 8000, 120000000
    ch = cm.make_channel(cf + 100000, cf, sr, n, w, squelch_dbfs=-40.0, afc=afc)
    ch.offset_hz = 100000.0 + shift_bins * (sr / n)
    cfg = cm.Config(fft_size=n, wave_rate=w, devices=[cm.Device(sample_rate=sr, sfmt=cm.SFMT_U8, centerfreq=cf, channels=[ch])])
    raws = [wl.synth_iq(cfg, 0, wl.samples_for_batches(cfg, 0, 5), key_on_s=0.25, key_off_s=0.15, amplitude=0.3)]
    ores, oorc = op.run_oracle(cfg, raws)
    gres, geng = lib.demodulate_all(cfg, raws, max_batches_per_run=1)
    want = ord('>') if shift_bins < 0 else ord('<')
    assert np.any(ores[0][2] == ord('>')) or np.any(ores[0][2] == ord('<')), "oracle AFC never moved: case is not exercising AFC"
    r = parity.strict(gres[0], ores[0])
    assert r["ok"], r
    assert geng.stats(0, 0).bin == oorc.stats(0, 0).bin
    geng.close()
