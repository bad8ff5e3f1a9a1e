"""Debug aid for the GPU-summed mixers behind the C++ host adapter: runs cfg4 through host.run_host_pipeline() and
least-squares fits every delivered mixer batch onto the per-device oracle outputs, so a missing or mis-aligned input shows up as a
missing / shifted weight.  Run on the GPU box:  python tools/mixdbg.py"""
import sys, os
ROOT = "/root/repo"
sys.path[:0] = [os.path.join(ROOT, 'rtlsdr-airband_b200', 'py'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests'), ROOT]
import numpy as np
import oracle_py as op
import parity
from airband_b200 import host, lib
from airband_b200 import workloads as wl
cfg = wl.cfg4()
nb = 5
raws = [wl.synth_iq(cfg, d, wl.samples_for_batches(cfg, d, nb), key_on_s=0.2, key_off_s=0.1) for d in range(len(cfg.devices))]
mixers = [[(d, m, 1.0 + 0.25 * d, (-0.5 if (m == 1 and d == 0) else 0.0)) for d in range(len(cfg.devices))] for m in range(4)]
ores, _ = op.run_oracle(cfg, raws)
got = []
hres = host.run_host_pipeline(cfg, raws, mixers=mixers, mixer_out=got)
B = cfg.wave_batch
for m in range(4):
    g = got[m]
    print("mixer", m, "batches", g["left"].shape, "overruns", g["overruns"], "axc", bytes(g["axc"].astype(np.uint8)))
    for b in range(g["left"].shape[0]):
        # least squares of got onto the per-device oracle waveouts of every batch
        cols, names = [], []
        for d in range(4):
            for bb in range(nb):
                cols.append(ores[d][0][m][bb * B:(bb + 1) * B].astype(np.float64)); names.append((d, bb))
        A = np.stack(cols, 1)
        x, *_ = np.linalg.lstsq(A, g["left"][b].astype(np.float64), rcond=None)
        sig = [(names[i], round(float(x[i]), 3)) for i in range(len(x)) if abs(x[i]) > 0.05]
        print("  batch", b, "weights", sig)
