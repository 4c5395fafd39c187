"""BASELINE.json configs[4] on the GPU: long-tap FIR stress (255-tap Kaiser designs for all three stages, steeper
than fir1cpp.C:8 / fir2cpp.C:22 / fir3cpp.h:16) through the long-tap path: one kernel per stage, stages 1 and 2 on the tensor
cores by default (fir_long_tc.cu: tcgen05 3xTF32 Toeplitz GEMM), stage 3 and the NVX_LONG_TC=0 variants register-tiled on
CUDA cores (fir_long.cu).

The unmodified reference cannot run other tap sets, so parity is against the C restatement of the three stages
(oracle/navtex_oracle.c, h1/h2/h3 parameters), which is pinned to the compiled reference at the default taps."""
import numpy as np
import pytest
from scipy import signal

import oracle_lib as ol
from navtex_b200 import engine, synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5


def designs(n1=255, n2=255, n3=255):
    h1 = signal.firwin(n1, 20000, window=("kaiser", 8.0), fs=252000)
    h2 = signal.firwin(n2, 2000, window=("kaiser", 8.0), fs=63000)
    h3 = signal.firwin(n3, 250, window=("kaiser", 7.0), fs=9000)
    return h1, h2, h3


def _capture(text, offset, seconds, snr, seed):
    em = synth.Emission(text, offset, start_s=0.3, n_phasing=18, n_tail=5)
    return synth.quantise_s16(synth.fsk_iq([em], seconds, snr_db=snr, seed=seed))


# (255, ...), (101, ...), (37, 47, 500): long-tap path; (61, 75, 111), (45, 47, 90): the fused kernel's medium class;
# (21, 31, 51): shorter than the reference, zero-padded into the reference class
# (383, 511, 71) and (516, 900, 71): streaming kernel with three live tiles in both stages; (255, 930, 71): stage 2 beyond the
# streaming kernel's 903 taps on the tile-at-a-time tensor-core kernel (and therefore the NCO mix in the stage-1 epilogue)
@pytest.mark.parametrize("lengths", [(255, 255, 255), (101, 383, 129), (37, 47, 500), (61, 75, 111), (45, 47, 90), (21, 31, 51), (383, 511, 71),
                                     (516, 900, 71), (255, 930, 71)])
def test_long_taps_against_restated_oracle(lengths):
    taps = designs(*lengths)
    seconds = 11.0
    n = int(seconds * 252000)
    rng = np.random.default_rng(21)
    iqs, want = [], []
    for s in range(3):
        text, bbbb = synth.random_message(rng, n_lines=1, words_per_line=3)
        occ = s % 2
        iqs.append(_capture(text, 14000.0 if occ == 0 else -14000.0, seconds, snr=-6.0, seed=400 + s))
        want.append((518 if occ == 0 else 490, bbbb, text))
    x = np.stack([iq.reshape(-1, 2) for iq in iqs])                 # int16: the long path converts while staging
    eng = engine.Engine(3, n, keep_bits=True, taps=taps)
    eng.push_host(np.ascontiguousarray(x))
    msgs = eng.poll_messages()
    y3 = eng.read_y3()
    for s in range(3):
        o = ol.run_oracle(iqs[s], h1=taps[0], h2=taps[1], h3=taps[2])
        scale = max(np.abs(o.y3["518"]).max(), np.abs(o.y3["490"]).max())
        for c, tag in enumerate(ol.CHANNELS):
            err = np.abs(y3[s, c].astype(np.complex128) - o.y3[tag]).max() / scale
            assert err <= REL_TOL, (lengths, s, tag, err)
        occ = s % 2
        bits, _ = eng.read_bits(s, occ)
        assert bits == o.bits[ol.CHANNELS[occ]]
        assert eng.read_events(s, occ) == o.events[ol.CHANNELS[occ]]
        assert o.messages == [want[s]]
        assert [m[1:] for m in msgs if m[0] == s] == [want[s]]
    eng.close()


@pytest.fixture
def cuda_core_stage1(monkeypatch):
    """Stage 1 of the long-tap path on the CUDA-core kernel (the engine reads NVX_LONG_TC when it is created; by default
    stage 1 runs on the tensor cores, fir_long_tc.cu)."""
    monkeypatch.setenv("NVX_LONG_TC", "0")


@pytest.mark.parametrize("lengths", [(255, 255, 255), (101, 383, 129)])
def test_long_taps_cuda_core_stage1(cuda_core_stage1, lengths):
    test_long_taps_against_restated_oracle(lengths)


@pytest.mark.parametrize("lengths", [(255, 255, 255), (127, 255, 90), (511, 255, 255), (511, 511, 90), (1000, 47, 71)])
def test_long_taps_tensor_core_stage1(monkeypatch, lengths):
    """tcgen05 3xTF32 Toeplitz GEMM for stage 1 (the default; 1000 taps loads the band matrix in many TMA boxes): same 1e-5 bar as the
    oracle tests above; here a ragged stream count (130 = one full 128-row tile + 2), float input, and blockings that move the
    tile boundaries -- equal to rounding, not bit for bit -- and the CUDA-core stage 1 on the same input."""
    taps = designs(*lengths)
    n = 252000 * 2
    rng = np.random.default_rng(27)
    x = np.rint(rng.normal(0, 3000, size=(130, n, 2))).astype(np.float32)
    ref = engine.Engine(130, n, taps=taps)
    ref.push_host(x)
    want = ref.read_y3()
    ref.close()
    scale = np.abs(want).max()
    for blk in (280 * 333, 2520 * 40):            # blockings move the tile boundaries: equal to rounding, not bit for bit
        eng = engine.Engine(130, blk, taps=taps)
        ys = []
        for a in range(0, n // blk * blk, blk):
            eng.push_host(np.ascontiguousarray(x[:, a:a + blk]))
            ys.append(eng.read_y3())
        eng.close()
        y = np.concatenate(ys, axis=2)
        assert np.abs(y - want[:, :, : y.shape[2]]).max() <= REL_TOL * scale, blk
    monkeypatch.setenv("NVX_LONG_TC", "0")        # and against the CUDA-core stage 1 on the same input
    cc = engine.Engine(130, n, taps=taps)
    cc.push_host(x)
    assert np.abs(cc.read_y3() - want).max() <= REL_TOL * scale
    cc.close()


@pytest.mark.parametrize("lengths", [(255, 255, 255), (383, 511, 71)])
def test_mix_on_load_equals_the_stage1_epilogue_mix(monkeypatch, lengths):
    """With the reference offsets the streaming tensor-core stage 2 rotates each sample by its channel's NCO phase while it
    converts it ("mix on load", the default: stage 1 writes ONE un-mixed 63 kHz row per stream); NVX_LONG_MIX=stage1 keeps the
    rotation in the stage-1 epilogue (one row per channel).  Same FP32 products, same MMA order per row: the 900 Hz samples are
    bit-identical -- over two pushes (carried, un-mixed history), a ragged stream count (130 = two 64-stream row blocks + 2) and
    a block that does not start at tick 0.  NVX_LONG_TC=2 puts the CUDA-core stage 1 (plain rows) in front of the same stage 2."""
    taps = designs(*lengths)
    blk = 2520 * 70
    rng = np.random.default_rng(29)
    x = np.rint(rng.normal(0, 3000, size=(130, 2 * blk + 280 * 37, 2))).astype(np.float32)

    def run():
        eng = engine.Engine(130, blk, taps=taps)
        ys = []
        for a, b in ((0, 280 * 37), (280 * 37, 280 * 37 + blk), (280 * 37 + blk, 280 * 37 + 2 * blk)):
            eng.push_host(np.ascontiguousarray(x[:, a:b]))
            ys.append(eng.read_y3())
        st = eng.stats()
        eng.close()
        assert st.long_tc_fallbacks == 0
        return np.concatenate(ys, axis=2)

    on_load = run()
    monkeypatch.setenv("NVX_LONG_MIX", "stage1")
    in_stage1 = run()
    assert np.abs(in_stage1).max() > 0
    assert np.array_equal(on_load.view(np.uint64), in_stage1.view(np.uint64))
    monkeypatch.delenv("NVX_LONG_MIX")
    monkeypatch.setenv("NVX_LONG_TC", "2")
    cc1 = run()
    assert np.abs(cc1 - on_load).max() <= REL_TOL * np.abs(on_load).max()


@pytest.mark.parametrize("lengths", [(255, 255, 255), (61, 75, 111)])
def test_long_taps_blocking_and_format_invariance(cuda_core_stage1, lengths):
    """Histories are carried per stage (long path, CUDA-core kernels) / recomputed from a longer input tail (medium class):
    any blocking, float or int16 input, gives bit-identical 900 Hz samples."""
    taps = designs(*lengths)
    n = 252000 * 2
    rng = np.random.default_rng(23)
    x = np.rint(rng.normal(0, 3000, size=(2, n, 2))).astype(np.int16)
    one = engine.Engine(2, n, taps=taps)
    one.push_host(x.astype(np.float32))
    ref = one.read_y3()
    one.close()
    for blk in (280, 280 * 333, 2520 * 40):
        m = n if blk > 280 else 280 * 300
        eng = engine.Engine(2, blk, taps=taps)
        ys = []
        for a in range(0, m, blk):
            eng.push_host(np.ascontiguousarray(x[:, a:a + blk]))
            ys.append(eng.read_y3())
        eng.close()
        y = np.concatenate(ys, axis=2)
        assert np.array_equal(y.view(np.uint64), ref[:, :, : y.shape[2]].view(np.uint64)), blk


@pytest.mark.parametrize("lengths", [(255, 255, 255), (61, 75, 111)])
def test_long_taps_with_per_stream_nco(lengths):
    taps = designs(*lengths)
    seconds = 11.0
    n = int(seconds * 252000)
    rng = np.random.default_rng(25)
    text, bbbb = synth.random_message(rng, n_lines=1, words_per_line=3)
    iq = _capture(text, 9500.0, seconds, snr=-6.0, seed=500)
    eng = engine.Engine(1, n, taps=taps, nco_hz=[[9500.0, -4500.5]], stream_freq_tag=[[4209, 4195]])
    eng.push_host(np.ascontiguousarray(iq.reshape(1, -1, 2)))
    msgs = eng.poll_messages()
    y3 = eng.read_y3()
    o = ol.run_oracle(iq, h1=taps[0], h2=taps[1], h3=taps[2], nco_hz=(9500.0, -4500.5), freq_tag=(4209, 4195))
    scale = max(np.abs(o.y3["518"]).max(), np.abs(o.y3["490"]).max())
    for c, tag in enumerate(ol.CHANNELS):
        assert np.abs(y3[0, c].astype(np.complex128) - o.y3[tag]).max() <= REL_TOL * scale
    assert [m[1:] for m in msgs] == o.messages == [(4209, bbbb, text)]
    eng.close()
