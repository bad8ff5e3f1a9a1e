#!/usr/bin/env python
"""Generates tests/golden/*.npz.  Run in the authoring container, where /root/reference exists:

    python tests/golden/make_golden.py

For every case in tests/cases.py listed in GOLDEN it stores the exact raw I/Q bytes fed in and the outputs of
oracle/_ref/libairband_ref.so — i.e. the reference's OWN squelch.cpp / ctcss.cpp / filters.cpp (compiled in place
from /root/reference/src by oracle/Makefile) behind the restated demodulate() loop and the FP32 FFT stand-in.
(The reference's main translation unit cannot be built here: lame/shout/libconfig++/fftw3 are absent.)
The fixtures travel to the GPU box, where /root/reference does not exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "rtlsdr-airband_b200", "py")):
    sys.path.insert(0, p)

import oracle_py as op  # noqa: E402
from cases import CASES  # noqa: E402

GOLDEN = ["am_u8", "nfm_s16", "am_bw_f32", "s8_two_devices"]


def main():
    assert op.available("ref"), "build oracle/_ref first (make -C oracle)"
    for name in GOLDEN:
        cfg, raws = CASES[name]()
        res, o = op.run_oracle(cfg, raws, "ref")
        out = {}
        for d, (wo, iq, ax) in enumerate(res):
            out[f"raw{d}"] = raws[d]
            out[f"waveout{d}"] = wo
            out[f"iq_out{d}"] = iq
            out[f"axc{d}"] = ax
            st = []
            for c in range(wo.shape[0]):
                s = o.stats(d, c)
                st.append([s.open_count, s.flappy_count, s.ctcss_count, s.no_ctcss_count, s.active_counter])
            out[f"counts{d}"] = np.array(st, np.int64)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
