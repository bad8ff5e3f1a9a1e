"""ORACLE-SIDE CHECKER — TEST INFRASTRUCTURE ONLY (used by tests/ and by bench.py's `parity_spot` legs; never by the product).

Two comparisons of a GPU result with the CPU oracle's on the same bytes:

strict   the gate of BASELINE.md §3: per audio sample |gpu - oracle| <= 1e-4 * max(1, |gpu|, |oracle|) AND identical
         axcindicate per batch.  Used wherever the workload keeps signals clear of the squelch thresholds
         (manual -30 dBFS level, SURVEY.md §7 hard part 3).

relaxed  for `squelch_snr_threshold = 0` (config/noaa.conf:24; the throughput variant of BASELINE configs[2]): the squelch
         level equals the noise floor, so the compare at reference src/squelch.cpp:463 rides its threshold and a 1e-7
         difference in an FFT bin can move an open/close edge by a sample.  SURVEY.md §7.3 prescribes what is done here:
         state-transition indices are compared separately from audio (how many edges have no partner within a few
         samples), and audio is gated only over samples where BOTH sides are open and no disagreeing edge is recent
         (the demodulator's IIR state needs a few hundred samples to forget an edge that moved).
"""
from __future__ import annotations

import numpy as np

TOL = 1e-4


def gate(a, b) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float((np.abs(a - b) / np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))).max())


def strict(g, o, tol: float = TOL) -> dict:
    """g, o = (waveout[C, n], iq_out[C, n] or None, axc[nb, C]) of one device."""
    gw, _, ga = g
    ow, _, oa = o
    if gw.shape != ow.shape:
        return {"ok": False, "why": f"shape {gw.shape} vs {ow.shape}"}
    err = gate(gw, ow)
    same_axc = bool(np.array_equal(ga, oa))
    return {"ok": bool(err <= tol and same_axc), "max_err": err, "axc_equal": same_axc, "opened": int((np.asarray(oa) != ord(' ')).sum())}


def _edges(mask: np.ndarray) -> np.ndarray:
    return np.flatnonzero(np.diff(mask.astype(np.int8)) != 0) + 1


def relaxed(g, o, tol: float = TOL, edge_slack: int = 4, settle: int = 400) -> dict:
    """Transition indices and audio compared separately (see the module docstring).  Open/closed per sample is read off
    the audio itself: a closed squelch writes exact zeros (reference src/rtl_airband.cpp:613-619)."""
    gw, _, _ = g
    ow, _, _ = o
    if gw.shape != ow.shape:
        return {"ok": False, "why": f"shape {gw.shape} vs {ow.shape}"}
    n_edges = n_unmatched = 0
    n_cmp = n_bad = 0
    worst = 0.0
    for c in range(gw.shape[0]):
        gm, om = gw[c] != 0.0, ow[c] != 0.0
        ge, oe = _edges(gm), _edges(om)
        n_edges += len(oe)
        # edges of either side without a partner within edge_slack samples on the other side
        bad_at = []
        for a, b in ((ge, oe), (oe, ge)):
            if len(a) == 0:
                continue
            if len(b) == 0:
                bad_at.extend(a.tolist())
                continue
            pos = np.searchsorted(b, a)
            lo = np.abs(a - b[np.clip(pos - 1, 0, len(b) - 1)])
            hi = np.abs(a - b[np.clip(pos, 0, len(b) - 1)])
            bad_at.extend(a[np.minimum(lo, hi) > edge_slack].tolist())
        n_unmatched += len(bad_at)
        both = gm & om
        if bad_at:  # the state behind a moved edge needs time to settle: leave those stretches out of the audio gate
            dirty = np.zeros(gw.shape[1], bool)
            for e in bad_at:
                dirty[max(0, e - edge_slack):e + settle] = True
            both &= ~dirty
        if both.any():
            a, b = gw[c][both].astype(np.float64), ow[c][both].astype(np.float64)
            err = np.abs(a - b) / np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
            n_cmp += int(err.size)
            n_bad += int((err > tol).sum())
            worst = max(worst, float(err.max()))
    frac_bad = n_bad / max(n_cmp, 1)
    frac_unmatched = n_unmatched / max(2 * n_edges, 1)
    return {"ok": bool(frac_bad <= 1e-3 and frac_unmatched <= 0.02), "edges": int(n_edges), "edges_unmatched": int(n_unmatched),
            "audio_samples_compared": int(n_cmp), "audio_samples_outside_gate": int(n_bad), "max_err_compared": worst}


def mixer_reference(cfg, ores, mixers, n_batches):
    """What mixer.cpp:133-140,189-214 produces from the oracle's per-channel audio: ref[m][b] = (left[B], right[B], has_signal).
    mixers[m] = [(dev, chan, ampfactor, balance), ...]; inputs are added in input order with the reference's float arithmetic."""
    B = cfg.wave_batch
    ref = []
    for inputs in mixers:
        per_batch = []
        for b in range(n_batches):
            left = np.zeros(B, np.float32)
            right = np.zeros(B, np.float32)
            sig = False
            for (d, c, amp, bal) in inputs:
                wo, _, ax = ores[d]
                if ax[b, c] == ord(' '):
                    continue
                sig = True
                ampl, ampr = np.float32(min(1.0, 1.0 - bal)), np.float32(min(1.0, 1.0 + bal))
                x = wo[c, b * B:(b + 1) * B]
                left = (left + x * (np.float32(amp) * ampl)).astype(np.float32)
                right = (right + x * (np.float32(amp) * ampr)).astype(np.float32)
            per_batch.append((left, right, sig))
        ref.append(per_batch)
    return ref
