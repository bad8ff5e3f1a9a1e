"""CPU-only checks of the ingest side of the host adapter: the "pattern" input plugin
(host/input_pattern.cpp, shape of reference src/input-file.cpp) feeding an input ring through circbuffer_append()
(reference src/input-helpers.cpp:37-63) while a consumer drains it like demodulate() does."""
import numpy as np

from airband_b200 import config as cm
from airband_b200 import host


def _block(nbytes, seed=7):
    return np.random.default_rng(seed).integers(0, 256, nbytes, dtype=np.uint8)


def test_lossless_replay_is_byte_exact_across_many_ring_wraps():
    blk = _block(2 * 150001)                      # not a divisor of the 256 KiB ring: every wrap lands somewhere else
    bad, consumed, overflows = host.pattern_selftest(blk, cm.SFMT_U8, 2560000, 512, repeat=9, speedup=0.0)
    assert (bad, overflows) == (0, 0)
    assert consumed == 9 * blk.nbytes > 8 * 256 * 1024


def test_paced_source_keeps_real_time_and_does_not_overflow_a_fast_consumer():
    blk = _block(2 * 64000, seed=3)               # 25 ms of a 2.56 Msps U8 stream
    bad, consumed, overflows = host.pattern_selftest(blk, cm.SFMT_U8, 2560000, 512, repeat=8, speedup=4.0)
    assert (bad, overflows) == (0, 0) and consumed == 8 * blk.nbytes


def test_slow_consumer_overflows_like_a_live_sdr():
    blk = _block(4 * 50000, seed=5)               # S16: 4 bytes per complex sample
    bad, consumed, overflows = host.pattern_selftest(blk, cm.SFMT_S16, 2560000, 1024, repeat=60, speedup=20.0, consumer_delay_us=30000)
    assert overflows >= 1                         # 204.8 MB/s into a 256 KiB ring drained every 30 ms
    assert bad == 0                               # everything read before the first overflow was still the right bytes
