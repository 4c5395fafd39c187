#!/usr/bin/env python3
"""Quick device-resident throughput probe of the two kernels (development aid; bench.py is the contract)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from navtex_b200 import engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=1024)
ap.add_argument("--super", type=int, default=9250, help="900 Hz outputs per stream per block")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--timing", type=int, default=2)
ap.add_argument("--s16", action="store_true", help="int16 I,Q input (fused ingest)")
ap.add_argument("--nco", action="store_true", help="general per-stream NCO variant")
ap.add_argument("--taps", type=int, default=0, help="N > 0: N-tap Kaiser designs for all three stages (long-tap path)")
a = ap.parse_args()
n = a.super * 280
x = torch.empty((a.streams, n, 2), dtype=torch.float32, device="cuda")
for s0 in range(0, a.streams, 64):
    x[s0:s0 + 64].normal_(0, 3000).round_()
if a.s16:
    x = x.to(torch.int16)
torch.cuda.synchronize()
nco = [[14000.0, -14000.0]] * a.streams if a.nco else None
taps = None
if a.taps:
    from scipy import signal
    taps = (signal.firwin(a.taps, 20000, window=("kaiser", 8.0), fs=252000), signal.firwin(a.taps, 2000, window=("kaiser", 8.0), fs=63000),
            signal.firwin(a.taps, 250, window=("kaiser", 7.0), fs=9000))
eng = engine.Engine(a.streams, n, nco_hz=nco, taps=taps)
eng.enable_timing(a.timing)
for _ in range(2):
    eng.push_device(x.data_ptr(), n, s16=a.s16)
eng.sync()
eng.stats()
t0 = time.time()
for _ in range(a.steps):
    eng.push_device(x.data_ptr(), n, s16=a.s16)
eng.sync()
wall = time.time() - t0
st = eng.stats()
tot = a.streams * n * a.steps
print(f"streams={a.streams} n={n} steps={a.steps} wall={wall*1e3/a.steps:.3f} ms/step "
      f"cascade={st.cascade_ms/a.steps:.3f} ms demod={st.demod_ms/a.steps:.3f} ms")
print("demod stages (ms/step): " + " ".join(f"{nm}={st.demod_stage_ms[k]/a.steps:.3f}" for k, nm in enumerate(("angle", "sum", "carry", "clock", "decide", "fsm"))))
if a.taps:
    flop = 4.0 * a.taps * (1 / 4 + 2 / 28 + 2 / 280) + 6 / 4      # real FMAs x 2 per complex-by-real tap, + mix
    print(f"long taps {a.taps}: {flop:.0f} flop/sample -> {tot*flop/st.cascade_ms/1e9:.1f} TFLOP/s FP32 in the three stage kernels")
print(f"cascade: {tot/st.cascade_ms/1e6:.1f} Gsamples/s = {tot*(4.06 if a.s16 else 8.06)/st.cascade_ms/1e6:.0f} GB/s algorithmic; "
      f"whole step (wall): {tot/wall/1e9:.1f} Gsamples/s")
