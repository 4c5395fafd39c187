"""Deterministic synthetic captures shared by the golden-fixture generator and the tests."""
from __future__ import annotations

import hashlib

import numpy as np

from navtex_b200 import synth

# name -> (builder kwargs).  Durations are multiples of 10 ms so that n % 2520 == 0.
CASES = {
    # short bulletin on 518 kHz, light noise
    "clean518": dict(text="ZCZC PA12\nGALE WARNING 8\nNNNN\n", offset=14000.0, duration=9.0, snr=10.0, seed=11, phasing=16),
    # same on 490 kHz, -12 dB full-band SNR
    "noisy490": dict(text="ZCZC QB07\nICE REPORT NIL\nNNNN\n", offset=-14000.0, duration=9.5, snr=-12.0, seed=12, phasing=18),
    # weak signal at the edge of the decode range: garbled characters, '*' marks, possible abort
    "weak518": dict(text="ZCZC LJ75\nSAR EXERCISE AREA CLOSED\nNNNN\n", offset=14000.0, duration=11.0, snr=-23.0, seed=13, phasing=20),
    # transmitter drops out mid-bulletin: error window overflows, partial message stored by message_abort
    "dropout": dict(text="ZCZC MK33\nDRIFTING CONTAINER REPORTED\nKEEP SHARP LOOKOUT\nNNNN\n", offset=14000.0, duration=12.0,
                    snr=-6.0, seed=15, phasing=16, stop=5.5),
    # noise only: free-running bit sync, random bits
    "noise": dict(text=None, offset=14000.0, duration=4.0, snr=0.0, seed=14, phasing=0),
    # both channels busy at once with different bulletins (nav_sched.C:10-16: two independent decoder chains behind one stage 1)
    "both": dict(text="ZCZC EA44\nNAVAREA ONE 123\nNNNN\n", offset=14000.0, duration=12.0, snr=-4.0, seed=16, phasing=16,
                 second=dict(text="ZCZC SL09\nWX FCST NIL\nNNNN\n", offset=-14000.0, start=0.9, phasing=20)),
    # figures shift, punctuation and digits (LTRS / FIGS state, nav_b_sm.C:100-145)
    "figures": dict(text="ZCZC BD57\n51-23.4N 002-45.6E (WIDE BERTH) 7/8 = 0.875?\nQRT: +12,5 KTS.\nNNNN\n", offset=-14000.0, duration=19.0, snr=0.0,
                    seed=17, phasing=14),
    # two emissions on one channel, 12.5 s apart: end of emission, detector hold-off (1100 bits, nav_b_sm.h:52) and re-phasing
    "twice": dict(text="ZCZC GA01\nFIRST\nNNNN\n", offset=14000.0, duration=31.0, snr=-2.0, seed=18, phasing=14,
                  second=dict(text="ZCZC GB02\nSECOND\nNNNN\n", offset=14000.0, start=16.0, phasing=14)),
}


def build(name: str) -> np.ndarray:
    """Interleaved int16 I,Q capture of a case."""
    c = CASES[name]
    if c["text"] is None:
        rng = np.random.default_rng(c["seed"])
        n = int(round(c["duration"] * synth.FS))
        x = 2000.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        return synth.quantise_s16(x)
    ems = [synth.Emission(c["text"], c["offset"], start_s=0.25, n_phasing=c["phasing"], n_tail=5, stop_s=c.get("stop"))]
    if "second" in c:
        d = c["second"]
        ems.append(synth.Emission(d["text"], d["offset"], start_s=d["start"], n_phasing=d["phasing"], n_tail=5))
    x = synth.fsk_iq(ems, c["duration"], snr_db=c["snr"], seed=c["seed"])
    return synth.quantise_s16(x)


def digest(iq: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(iq).tobytes()).hexdigest()
