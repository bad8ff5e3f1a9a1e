"""The steps either side of the path through the C++ host adapter (rtlsdr-airband_b200/host), driven the way the reference's
threads drive them: mixer sums handed to the output thread through mixer_t.channel with the CH_DIRTY -> CH_WORKING -> CH_READY
handshake (reference src/mixer.cpp:157-261 producer, src/output.cpp:888-896 consumer), the raw-I/Q file format
(src/output.cpp:519-522), disable_device_outputs() for a finished receiver (src/rtl_airband.cpp:386) and the dBFS levels of the
stats file / TUI (src/util.cpp:169-180)."""
import numpy as np
import pytest

import oracle_py as op
import parity
from airband_b200 import host, lib
from airband_b200 import workloads as wl
from cases import CASES

pytestmark = pytest.mark.gpu


def test_mixer_sums_reach_the_output_thread_through_the_channel_state_handshake():
    cfg = wl.cfg4()
    nb = 5
    raws = [wl.synth_iq(cfg, d, wl.samples_for_batches(cfg, d, nb), key_on_s=0.2, key_off_s=0.1) for d in range(len(cfg.devices))]
    mixers = [[(d, m, 1.0 + 0.25 * d, (-0.5 if (m == 1 and d == 0) else 0.0)) for d in range(len(cfg.devices))] for m in range(4)]
    ores, _ = op.run_oracle(cfg, raws)
    got = []
    hres = host.run_host_pipeline(cfg, raws, mixers=mixers, mixer_out=got)
    for d in range(len(raws)):
        assert parity.strict((hres[d][0], None, hres[d][2]), ores[d])["ok"]
    ref = parity.mixer_reference(cfg, ores, mixers, nb)
    assert len(got) == 4
    for m in range(4):
        g = got[m]
        assert g["is_gpu"] and g["overruns"] == 0
        assert g["left"].shape == (nb, cfg.wave_batch)
        for b in range(nb):
            left, right, sig = ref[m][b]
            assert bool(g["axc"][b] != ord(' ')) == sig
            assert parity.gate(g["left"][b], left) <= parity.TOL
            if m == 1:  # the stereo mixer (one input panned): waveout_r is delivered too
                assert parity.gate(g["right"][b], right) <= parity.TOL
    assert any(ref[m][b][2] for m in range(4) for b in range(nb)), "no mixer batch carried signal"


def test_rawfile_output_is_interleaved_float32_iq(tmp_path):
    cfg, raws = CASES["am_u8"](n_batches=4)  # channel 1 has a rawfile output (has_iq_outputs)
    path = tmp_path / "chan1.cf32"
    hres = host.run_host_pipeline(cfg, raws, rawfiles=[(0, 1, path)])
    gw, gi, ga, info = hres[0]
    data = np.fromfile(path, dtype=np.float32)
    B = cfg.wave_batch
    assert data.size == 4 * 2 * B                      # 2 * sizeof(float) * WAVE_BATCH bytes per batch (output.cpp:519-520)
    assert data.tobytes() == np.ascontiguousarray(gi[1]).view(np.float32).tobytes()   # byte for byte what channel_t.iq_out held
    ores, _ = op.run_oracle(cfg, raws)
    oi = ores[0][1][1]
    assert np.abs(oi).max() > 0.1
    assert parity.gate(data[0::2], oi.real) <= parity.TOL and parity.gate(data[1::2], oi.imag) <= parity.TOL


def test_finished_receiver_gets_its_outputs_disabled_once():
    cfg, raws = CASES["s8_two_devices"]()
    hres = host.run_host_pipeline(cfg, raws)
    for d in range(2):
        assert hres[d][3]["disable_device_outputs_calls"] == 1


def test_stats_levels_in_dbfs_match_level_to_dbfs():
    cfg, raws = CASES["am_u8"]()
    ores, oorc = op.run_oracle(cfg, raws)
    gres, geng = lib.demodulate_all(cfg, raws)
    L = op.lib()
    for c in range(2):
        gs, os_ = geng.stats(0, c), oorc.stats(0, c)
        for lvl, dbfs in ((os_.noise_level, gs.noise_level_dbfs), (os_.signal_level, gs.signal_level_dbfs), (os_.squelch_level, gs.squelch_level_dbfs)):
            want = L.abo_level_to_dbfs(lvl, cfg.fft_size)
            assert abs(dbfs - want) <= 1e-3, (c, dbfs, want)
            assert dbfs <= 0.0
    geng.close()
