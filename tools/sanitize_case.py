#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel variant, meant to run under compute-sanitizer (development aid)."""
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
medium_taps = (signal.firwin(61, 20000, fs=252000), signal.firwin(75, 2000, fs=63000), signal.firwin(111, 300, fs=9000))
for name, kw, dtype in (("f32", {}, np.float32), ("s16", {}, np.int16), ("nco f32", dict(nco_hz=nco), np.float32),
                        ("nco s16", dict(nco_hz=nco), np.int16), ("taps f32", dict(taps=short_taps), np.float32),
                        ("taps s16 nco", dict(taps=short_taps, nco_hz=nco), np.int16), ("medium f32", dict(taps=medium_taps), np.float32), ("medium s16 nco", dict(taps=medium_taps, nco_hz=nco), np.int16),
                        ("long s16", dict(taps=long_taps), np.int16),
                        ("long f32 nco", dict(taps=long_taps, nco_hz=nco), np.float32)):
    eng = engine.Engine(S, n, keep_bits=True, **kw)
    for k in range(3):
        eng.push_host(np.ascontiguousarray(x[:, k * n:(k + 1) * n]).astype(dtype))
    eng.push_host(np.ascontiguousarray(x[:, :280]).astype(dtype))       # a one-superblock block
    msgs = eng.poll_messages()
    y = eng.read_y3()
    bits, _ = eng.read_bits(S - 1, 1)
    assert np.isfinite(y.view(np.float32)).all()
    eng.close()
    print(name, "ok", len(msgs), len(bits), flush=True)
print("all variants ok")
