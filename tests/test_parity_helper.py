"""CPU test of tests/parity.py, the comparison behind every "1e-5 relative" claim: denominators, bit-decision bookkeeping and the
timing-slip attribution, on oracle output perturbed in known ways (no GPU involved)."""
import numpy as np

import cases
import oracle_lib as ol
import parity


def _as_gpu(o, scale_err=0.0, flip=None, seed=0):
    """Dress an oracle result up as GPU taps: complex64 samples (+ optional relative perturbation), bits with optional flips."""
    rng = np.random.default_rng(seed)
    y3 = np.stack([o.y3[t] for t in ol.CHANNELS]).astype(np.complex64)
    if scale_err:
        peak = np.abs(y3).max()
        y3 = y3 + (scale_err * peak * (rng.standard_normal(y3.shape) + 1j * rng.standard_normal(y3.shape)) / 4).astype(np.complex64)
    bits = {c: bytearray(o.bits[t]) for c, t in enumerate(ol.CHANNELS)}
    for c, k in (flip or []):
        bits[c][k] = ord("B") if bits[c][k] == ord("Y") else ord("Y")
    return y3, {c: bytes(b) for c, b in bits.items()}, {c: o.disc[t].copy() for c, t in enumerate(ol.CHANNELS)}


def test_identical_taps_pass_and_float32_rounding_is_measured():
    o = ol.run_oracle(cases.build("clean518"))
    y3, bits, disc = _as_gpu(o)
    cmp = parity.compare_stream(y3, bits, disc, o, {"518"})
    parity.assert_stream(cmp, {"518"})
    assert 0 < cmp["y3_rel_pair_peak"] < 1e-7                       # complex64 rounding of FP64 samples
    assert cmp["bit_mismatches"] == {"518": 0, "490": 0} and cmp["disc_rel"]["518"] == 0.0
    assert cmp["y3_rel_own_rms"]["518"] >= cmp["y3_rel_pair_peak"]   # the RMS of a channel is below the pair peak: the tighter bar


def test_errors_beyond_the_bar_and_flipped_decisions_fail():
    import pytest

    o = ol.run_oracle(cases.build("clean518"))
    y3, bits, disc = _as_gpu(o, scale_err=1e-4)
    with pytest.raises(AssertionError):
        parity.assert_stream(parity.compare_stream(y3, bits, disc, o, {"518"}), {"518"})
    y3, bits, disc = _as_gpu(o, flip=[(0, 100)])
    cmp = parity.compare_stream(y3, bits, disc, o, {"518"})
    assert cmp["bit_mismatches"]["518"] == 1 and len(cmp["mismatch_margins"]["518"]) == 1
    with pytest.raises(AssertionError):
        parity.assert_stream(cmp, {"518"})


def test_empty_channel_differences_are_reported_not_asserted():
    o = ol.run_oracle(cases.build("clean518"))
    y3, bits, disc = _as_gpu(o, flip=[(1, 50), (1, 51), (1, 52), (1, 400)])
    cmp = parity.compare_stream(y3, bits, disc, o, {"518"})
    parity.assert_stream(cmp, {"518"})                               # the empty channel's decisions are counted, not required equal
    assert cmp["bit_mismatches"]["490"] == 4 and len(cmp["mismatch_margins"]["490"]) == 4
    assert all(0.0 <= m <= 1.0 for m in cmp["mismatch_margins"]["490"])
    # two runs of differences, each attributed to the tightest arg max of the 64 evaluations before it
    assert len(cmp["timing_slip_pick_margins"]["490"]) == 2 and all(0.0 <= m <= 1.0 for m in cmp["timing_slip_pick_margins"]["490"])
    s = parity.summarise([cmp], [{"518"}])
    assert s["bit_mismatches_occupied"] == 0 and s["empty_channel_bit_mismatches"] == 4 and s["empty_channel_bits_compared"] == len(o.bits["490"])
