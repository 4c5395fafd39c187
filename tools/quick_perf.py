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
a = ap.parse_args()
n = a.super * 280
x = torch.empty((a.streams, n, 2), dtype=torch.float32, device="cuda")
for s0 in range(0, a.streams, 64):
    x[s0:s0 + 64].normal_(0, 3000).round_()
if a.s16:
    x = x.to(torch.int16)
torch.cuda.synchronize()
nco = [[14000.0, -14000.0]] * a.streams if a.nco else None
eng = engine.Engine(a.streams, n, nco_hz=nco)
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
print(f"cascade: {tot/st.cascade_ms/1e6:.1f} Gsamples/s = {tot*(4.06 if a.s16 else 8.06)/st.cascade_ms/1e6:.0f} GB/s algorithmic; "
      f"whole step (wall): {tot/wall/1e9:.1f} Gsamples/s")
