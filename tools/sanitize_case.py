#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel variant (development aid): fused cascade in both tap classes and input formats,
table / per-stream NCO, 1 / 3 / 4 / 7 channels per capture, long-tap path with the streaming and the tile-at-a-time tensor-core
kernels.  Written to run under compute-sanitizer; that tool is closed on this GPU pool, so it serves as a plain smoke run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from navtex_b200 import engine  # noqa: E402
from scipy import signal  # noqa: E402

rng = np.random.default_rng(0)
S, n = 35, 280 * 90
x = np.rint(rng.normal(0, 3000, size=(S, 3 * n, 2))).astype(np.int16)
nco = np.tile(np.array([[9500.0, -4500.5]]), (S, 1))
long_taps = (signal.firwin(255, 20000, fs=252000), signal.firwin(129, 2000, fs=63000), signal.firwin(300, 250, fs=9000))
short_taps = (signal.firwin(37, 20000, fs=252000), signal.firwin(47, 2000, fs=63000), signal.firwin(71, 300, fs=9000))
long511 = (signal.firwin(511, 20000, fs=252000), signal.firwin(511, 2000, fs=63000), signal.firwin(71, 250, fs=9000))
offs = [14000.0, -14000.0, 7000.0, -21000.0, 21000.5, -7000.5, 0.0]
nco1, nco3, nco4, nco7 = (np.tile(np.array([offs[:c]]), (S, 1)) for c in (1, 3, 4, 7))
medium_taps = (signal.firwin(61, 20000, fs=252000), signal.firwin(75, 2000, fs=63000), signal.firwin(111, 300, fs=9000))
for name, kw, dtype in (("f32", {}, np.float32), ("s16", {}, np.int16), ("nco f32", dict(nco_hz=nco), np.float32),
                        ("nco s16", dict(nco_hz=nco), np.int16), ("taps f32", dict(taps=short_taps), np.float32),
                        ("taps s16 nco", dict(taps=short_taps, nco_hz=nco), np.int16), ("medium f32", dict(taps=medium_taps), np.float32), ("medium s16 nco", dict(taps=medium_taps, nco_hz=nco), np.int16),
                        ("long s16", dict(taps=long_taps), np.int16),
                        ("long f32 nco", dict(taps=long_taps, nco_hz=nco), np.float32),
                        ("long 511 f32 (tile-at-a-time kernel)", dict(taps=long511), np.float32),
                        ("3 channels s16", dict(nco_hz=nco3, n_channels=3), np.int16), ("4 channels f32", dict(nco_hz=nco4, n_channels=4), np.float32),
                        ("7 channels f32 (two passes)", dict(nco_hz=nco7, n_channels=7), np.float32), ("1 channel s16", dict(nco_hz=nco1, n_channels=1), np.int16)):
    eng = engine.Engine(S, n, keep_bits=True, **kw)
    for k in range(3):
        eng.push_host(np.ascontiguousarray(x[:, k * n:(k + 1) * n]).astype(dtype))
    eng.push_host(np.ascontiguousarray(x[:, :280]).astype(dtype))       # a one-superblock block
    msgs = eng.poll_messages()
    y = eng.read_y3()
    bits, _ = eng.read_bits(S - 1, min(1, eng.C - 1))
    assert np.isfinite(y.view(np.float32)).all()
    eng.close()
    print(name, "ok", len(msgs), len(bits), flush=True)
print("all variants ok")
