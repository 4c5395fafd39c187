"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, the host-side message assembly matches the oracle, and compute entry points fail loudly
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cases
import oracle_lib as ol
from navtex_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:nvx_|init_fir|sample_in_|navtex_compat_)\w*)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = engine.load_library()
    names = [n for n in declared("navtex_b200.h") if n != "nvx_message_cb"]
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert set(engine.EXPORTS) <= set(names)


def test_compat_library_exports_reference_symbols():
    path = os.path.join(ROOT, "navtex_b200", "libnavtex_compat.so")
    L = C.CDLL(path)
    for n in ("init_fir_filter1", "sample_in_1", "init_fir2_wrapper", "navtex_compat_flush", "navtex_compat_set_sink"):
        assert hasattr(L, n), n
    # the C++-linkage half of the seam, under the reference's own mangled names (fir2cpp.h:3-6, fir3cpp.h:98-100,
    # decoder.h:84-85, nav_b_sm.h:126-127): what fir1cpp.o / nav_sched.o of the reference would bind
    for n in ("_Z16init_fir_filter2P11fir_filter3S0_", "_Z11sample_in_2dd", "_Z8fir_in_2dd", "_Z12fir_in_2_490dd",
              "_ZN11fir_filter3C1EP7decoder", "_ZN11fir_filter39sample_inEdd", "_ZN7decoderC1EP18byte_state_machine",
              "_ZN7decoder9sample_inEdd", "_ZN18byte_state_machineC1Ej", "_ZN18byte_state_machine11receive_bitEc"):
        assert hasattr(L, n), n


REF_DIR = "/root/reference/receiver"


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference tree not mounted (GPU box): the prebuilt binary is used there")
def test_reference_nav_sched_and_wav_compile_unmodified_against_the_seam():
    """The reference's OWN receiver/nav_sched.C (includes fir2cpp.h fir3cpp.h decoder.h nav_b_sm.h nav_sched.h) compiles
    unmodified against include/compat/, its own receiver/wav.c beside it, and both link with the WAV -> sample_in_1 driver
    against libnavtex_compat.so in place of the reference's DSP objects (oracle/Makefile, target compat_wav_host)."""
    import subprocess

    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/compat_wav_host"], check=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "compat_wav_host")
    nm = subprocess.run(["nm", exe], capture_output=True, text=True, check=True).stdout
    # nav_sched.C's own wiring is in the binary and calls OUR init_fir_filter2; the DSP symbols are left to the library
    assert re.search(r"\bT init_fir2_wrapper\b", nm) and re.search(r"\bU _Z16init_fir_filter2P11fir_filter3S0_", nm)
    assert re.search(r"\bU sample_in_1\b", nm) and re.search(r"\bT wav_read\b", nm) and re.search(r"\bT add_message\b", nm)
    for sym in ("_ZN18byte_state_machineC1Ej", "_ZN7decoderC1EP18byte_state_machine", "_ZN11fir_filter3C1EP7decoder"):
        assert re.search(r"\bU " + sym, nm), sym
    import torch
    if not torch.cuda.is_available():      # and without a GPU it fails loudly instead of decoding on the CPU
        from navtex_b200 import synth
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            wav = os.path.join(td, "c.wav")
            synth.write_wav(wav, cases.build("clean518"))
            cp = subprocess.run([exe, wav], capture_output=True, text=True)
            assert cp.returncode != 0 and cp.stdout == "" and "no CPU path" in cp.stderr


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.NvxError, match="no usable CUDA device|CUDA"):
        engine.Engine(1, 2800)


@pytest.mark.parametrize("name", ["clean518", "weak518", "dropout"])
def test_host_assembler_matches_oracle(name):
    """Feeding the oracle's event stream through the product's host assembler reproduces the oracle's messages."""
    iq = cases.build(name)
    r = ol.run_oracle(iq, record_taps=False)
    got = engine.host_assemble(r.events["518"], stream=7, freq=518)
    want = [(7, f, b, t) for f, b, t in r.messages if f == 518]
    assert got == want


def test_host_assembler_quirks():
    # a second ZCZC before NNNN keeps the first B1B2B3B4 (strncat onto a non-empty buffer, nav_b_sm.C:74-76)
    ev = b"ZCZC AB12\nTEXT\nZCZC CD34\nMORE\nNNNN\n"
    assert engine.host_assemble(ev) == [(0, 518, "AB12", "ZCZC CD34\nMORE\nNNNN\n")]
    # one garbled character in the framing words is tolerated (regex alternatives)
    assert engine.host_assemble(b"Z*ZC EF56\nX\nN*NN\n") == [(0, 518, "EF56", "Z*ZC EF56\nX\nN*NN\n")]
    # abort stores the partial message, and nothing when none is in progress
    assert engine.host_assemble(b"ZCZC GH78\nPART\x18") == [(0, 518, "GH78", "ZCZC GH78\n")]
    assert engine.host_assemble(b"NOISE\x18NNNN\n") == []


def test_message_store_replaces_by_bbbb_and_purges(tmp_path):
    """message_store.c:59-97 semantics: add = delete-by-bbbb + insert with a UTC minute stamp; purge by age."""
    st = engine.Store()
    t0 = 1_700_000_000
    assert st.add(0, "PA12", "ZCZC PA12\nOLD\nNNNN\n", 518, when=t0) == 0
    assert st.add(0, "QB07", "ZCZC QB07\nICE\nNNNN\n", 490, when=t0 + 60) == 0
    assert st.add(0, "PA12", "ZCZC PA12\nNEW \"quoted\"\nNNNN\n", 518, when=t0 + 120) == 0      # replaces the first
    assert st.add(1, "PA12", "ZCZC PA12\nOTHER RADIO\nNNNN\n", 518, when=t0 + 180) == 0        # another stream: kept apart
    rows = st.rows()
    assert [(r[0], r[2]) for r in rows] == [(0, "QB07"), (0, "PA12"), (1, "PA12")]
    assert rows[1][4].startswith("ZCZC PA12\nNEW") and rows[1][3] == "2023-11-14 22:15"
    path = str(tmp_path / "messages.csv")
    st.dump_csv(path)
    lines = open(path).read().splitlines()
    assert lines[0] == "id,stream,freq,bbbb,timestamp,message" and len(lines) == 4
    assert '""quoted""' in lines[2] and "\\n" in lines[2]
    assert st.purge(now=t0 + 100 + 72 * 3600) == 1 and len(st.rows()) == 2      # only the t0 + 60 row is older than 72 h
    st.close()


def build_example(name, out_dir):
    """Compile examples/<name> against the public headers and the in-tree libraries (plain gcc / g++, no nvcc)."""
    import subprocess

    src = os.path.join(ROOT, "examples", name)
    exe = os.path.join(str(out_dir), os.path.splitext(name)[0])
    lib_dir = os.path.join(ROOT, "navtex_b200")
    if name.endswith(".c"):
        cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), src, "-L" + lib_dir, "-lnavtex_b200"]
    else:
        cmd = ["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), src, "-L" + lib_dir, "-lnavtex_compat",
               "-lnavtex_b200"]
    subprocess.run(cmd + ["-Wl,-rpath," + lib_dir, "-o", exe], check=True)
    return exe


def test_example_hosts_compile_as_c_and_cpp(tmp_path):
    """navtex_b200.h is valid C11, navtex_compat.h valid C++; a host using only the reference's names links."""
    import subprocess
    import torch

    exe = build_example("wav_decode.c", tmp_path)
    build_example("relink_host.cpp", tmp_path)
    if not torch.cuda.is_available():
        from navtex_b200 import synth
        wav = str(tmp_path / "c.wav")
        synth.write_wav(wav, cases.build("clean518"))
        cp = subprocess.run([exe, wav], capture_output=True, text=True)
        assert cp.returncode == 1 and "no CPU path" in cp.stderr          # fails loudly, decodes nothing
        assert cp.stdout == ""
