"""Soak parity on the GPU: many streams with several emissions each (both channels, overlapping the 11 s phasing
hold-off or well clear of it, dropouts, weak and strong), pushed in ragged blocks -- every channel's character / line /
abort events and every add_message call identical to the CPU oracle's."""
import numpy as np
import pytest

import oracle_lib as ol
from navtex_b200 import engine, synth

pytestmark = pytest.mark.gpu


def _stream(rng, seconds):
    ems, t = [], 0.3 + 2.0 * rng.random()
    while True:
        text, _ = synth.random_message(rng, n_lines=int(rng.integers(1, 3)), words_per_line=int(rng.integers(2, 5)))
        n_ph = int(rng.integers(12, 40))
        dur = len(synth.message_bits(text, n_phasing=n_ph, n_tail=5)) / 100.0
        if t + dur + 0.5 > seconds:
            break
        stop = float(dur * rng.uniform(0.3, 0.8)) if rng.random() < 0.2 else None          # transmitter drops out
        ems.append(synth.Emission(text, 14000.0 if rng.random() < 0.5 else -14000.0, start_s=t, amplitude=float(rng.uniform(3000, 9000)),
                                  n_phasing=n_ph, n_tail=5, stop_s=stop))
        t += dur + float(rng.choice([0.4, 1.5, 4.0, 13.0]))                                # inside / outside the hold-off
    return ems


def test_many_emissions_ragged_blocks_match_oracle():
    S, seconds = 16, 45.0
    n = int(seconds * 252000)
    rng = np.random.default_rng(2024)
    iqs = []
    for s in range(S):
        x = synth.fsk_iq(_stream(rng, seconds), seconds, snr_db=float(rng.uniform(-22.0, -6.0)), seed=1000 + s)
        iqs.append(synth.quantise_s16(x))
    x = np.stack([iq.reshape(-1, 2) for iq in iqs])
    eng = engine.Engine(S, 280 * 4000)
    msgs, events = [], [[b"", b""] for _ in range(S)]
    pos = 0
    while pos < n:
        blk = min(n - pos, 280 * int(rng.integers(1, 4001)))
        eng.push_host(np.ascontiguousarray(x[:, pos:pos + blk]))
        msgs += eng.poll_messages()
        for s in range(S):
            for c in range(2):
                events[s][c] += eng.read_events(s, c)
        pos += blk
    eng.close()
    total = 0
    for s in range(S):
        o = ol.run_oracle(iqs[s], record_taps=False)
        for c, tag in enumerate(ol.CHANNELS):
            assert events[s][c] == o.events[tag], (s, tag)
        assert [m[1:] for m in msgs if m[0] == s] == o.messages, s
        total += len(o.messages)
    assert total >= S          # the sweep really decodes traffic (several messages per stream on average)
