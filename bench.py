#!/usr/bin/env python3
"""bench.py -- headline benchmark of the NAVTEX receive chain (BASELINE.json metric:
"IQ Msamples/s through FIR cascade + FSK demod at 1-8 B200; % HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the whole hot path (fused FIR cascade -> demod / bit-sync -> SITOR-B state
machine -> host message assembly) over one block of BASELINE.json configs[1]: 1024 synthetic IQ
streams per GPU (distinct SITOR-B bulletins, start times, channels and SNRs) x 10.28 s at 252 kS/s,
resident in HBM (21 GB per GPU: far larger than the 126 MB L2, so nothing is cache-warm between
steps).  Streams are independent: N GPUs = N x 1024 streams, no collective on the data path
(weak scaling); the only cross-rank traffic is the barrier and the max-over-ranks timing.

value  = samples processed by all ranks / max-over-ranks device time (CUDA events on the engine stream).
e2e    = same metric through nvx_engine_push_host_s16 (the reference's own int16 sample format) with
         pinned HOST buffers: the same captures pushed as consecutive 1.03 s blocks; H2D copy, int16->float
         conversion, all kernels, event download and host message assembly inside the timed region, and the
         decoded bulletins checked against what was transmitted.
--impl reference times the reference's own CPU chain (oracle/_ref/ref_chain, built unmodified from
the reference sources) on all host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAMS_PER_GPU = 1024
SUPER_PER_BLOCK = 9250              # 900 Hz outputs per stream per step -> n = 2,590,000 samples (10.28 s)
BLOCK = SUPER_PER_BLOCK * 280
E2E_CHUNKS = 10                     # the e2e arm pushes the same captures as 10 consecutive host blocks ...
E2E_BLOCK = BLOCK // E2E_CHUNKS     # ... of 259,000 samples (1.03 s) per stream: 1.06 GB of int16 per step
BYTES_PER_SAMPLE = 8.0 + 16.0 / 280 # algorithmic HBM bytes per input IQ sample: float2 in, 2 x float2 per 280 out
REF_CHAIN = os.path.join(ROOT, "oracle", "_ref", "ref_chain")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "cascade_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        # samples inside the timed region; a region shorter than the sampling period falls back to the samples
        # taken while the warm-up steps (same kernels, same load) were running just before it
        inside = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.05 and len(r) >= 7] or \
                 [r for ts, r in self.rows if t0 - 0.5 <= ts <= t1 + 0.15 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in inside)]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[0]) for r in inside if num(r[0]) is not None]
        pw = [num(r[2]) for r in inside if num(r[2]) is not None]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": num(inside[0][1]), "power_w_max": max(pw) if pw else None,
                "samples": len(inside), "reasons": reasons}


def bind_to_gpu_numa_node(local):
    """Run this rank (and so its pinned host buffers: first touch) on the CPUs of the NUMA node its GPU hangs off.
    Matters for the e2e arm at N > 1, where every GPU pulls 55 GB/s out of host memory.  No-op when the topology is not
    visible (containers without sysfs NUMA info)."""
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.startswith("0000"):
            bdf = bdf[4:]                      # nvidia-smi prints an 8-digit domain, sysfs uses 4
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        use = cpus & os.sched_getaffinity(0)
        if not use:
            return None
        os.sched_setaffinity(0, use)
        return node, len(use)
    except Exception:
        return None


def build_workload(torch, device, rank):
    """1024 distinct synthetic captures for this rank, generated on the device (untimed)."""
    import numpy as np
    from navtex_b200 import engine, synth

    S = STREAMS_PER_GPU
    rng = np.random.default_rng(518490 + rank)
    bits, off, start, amp, sigma, expect = [], [], [], [], [], []
    for s in range(S):
        while True:      # bulletins short enough to fit one 10.28 s block together with their start delay
            text, bbbb = synth.random_message(rng, n_lines=1, words_per_line=3)
            b = synth.message_bits(text, n_phasing=18, n_tail=5)
            if len(b) * 2520 + 0.6 * 252000 < BLOCK:
                break
        bits.append(b)
        ch = s % 2
        off.append(14000.0 if ch == 0 else -14000.0)
        start.append(0.05 + 0.4 * rng.random())
        amp.append(3000.0 + 6000.0 * rng.random())
        sigma.append(amp[-1] * 10 ** (rng.uniform(-6.0, 14.0) / 20) / np.sqrt(2))     # full-band SNR -14 .. +6 dB
        expect.append((s + rank * S, 518 if ch == 0 else 490, bbbb, text))
    x = torch.empty((S, BLOCK, 2), dtype=torch.float32, device=device)
    engine.synth_fill_device(device.index, x.data_ptr(), S, 0, BLOCK, bits, off, start, amp, sigma, seed=518490 + rank)
    return x, expect


def run_reference_cpu(samples_i16, passes_warm, passes_timed, cores):
    """Time the unmodified reference chain on `cores` host cores, one process (= one stream: its state is
    global) per core, `passes` back-to-back passes each.  Returns aggregate samples/s over the timed passes."""
    import numpy as np

    n_streams = len(samples_i16)
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for k, iq in enumerate(samples_i16):
            p = os.path.join(td, f"s{k}.s16")
            np.ascontiguousarray(iq, dtype=np.int16).tofile(p)
            paths.append(p)
        procs = [subprocess.Popen([REF_CHAIN, "--s16", paths[k % n_streams], "--passes", str(passes_warm + passes_timed)],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True) for k in range(cores)]
        outs = [p.communicate()[1] for p in procs]
    per_pass = []
    n = samples_i16[0].size // 2
    for o in outs:
        j = json.loads(o.strip().splitlines()[-1])
        per_pass.append(j["pass_s"][passes_warm:])
    # all processes run concurrently: pass k of the job ends when the slowest core finishes it
    step_s = [max(pp[k] for pp in per_pass) for k in range(passes_timed)]
    total = cores * n * passes_timed
    return total / sum(step_s), sum(step_s) / passes_timed, n


def impl_reference(args, rank, world):
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from navtex_b200 import synth

    cores = os.cpu_count() or 1
    rng = np.random.default_rng(518490)
    n_sample_streams = min(cores, 16)
    streams = []
    seconds = 4.0          # bounded sample: 4 s of each of up to 16 of the workload's streams
    for s in range(n_sample_streams):
        text, _ = synth.random_message(rng, n_lines=1, words_per_line=2)
        em = synth.Emission(text, 14000.0 if s % 2 == 0 else -14000.0, start_s=0.1, n_phasing=12, n_tail=4)
        streams.append(synth.quantise_s16(synth.fsk_iq([em], seconds, snr_db=float(rng.uniform(-14, 6)), seed=s)))
    if os.path.exists(REF_CHAIN):
        sps, step_s, n = run_reference_cpu(streams, args.warmup, args.steps, cores)
        kind = "reference"
    else:   # the compiled reference did not travel: time the C port instead
        import oracle_lib as ol
        t0 = time.time()
        for k in range(args.steps):
            ol.run_oracle(streams[k % len(streams)], record_taps=False)
        step_s = (time.time() - t0) / args.steps
        n, cores, kind = streams[0].size // 2, 1, "port"
        sps = n / step_s
    val = sps / 1e6
    sample = f"{cores} processes x {n} samples ({seconds:.0f} s of one workload stream each) per step"
    line = {
        "impl": "reference", "metric": "iq_msamples_per_s", "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: 1024 synthetic IQ streams/GPU, fused FIR cascade + FSK demod + bit-sync + SITOR-B (CPU chain on host cores, bounded sample)",
                   "streams_per_gpu": STREAMS_PER_GPU, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "realtime_streams": val / 0.252,
    }
    print(json.dumps(line), flush=True)


def run_config4(args, torch, dist, device, rank, world, local, barrier):
    """BASELINE.json configs[3]: 8192 streams per GPU (65536 over 8 GPUs) x 60 s of traffic, too large to be resident
    (124 GB per 8 s as float2): generated on the device block by block (counter-based generator keyed by
    (seed, stream, absolute sample): any block of any stream is reproducible) and decoded through the carried state.
    Generation is untimed: see the comment at the timed loop."""
    import numpy as np
    from navtex_b200 import engine, sharding, synth

    S = 8192
    block = 1125 * 280                               # 1.25 s per stream per block: 20.6 GB of float2 per GPU
    n_blocks = max(1, int(round(args.seconds * 252000 / block)))
    seconds = n_blocks * block / 252000
    rng = np.random.default_rng(65536 + rank)
    bits, off, start, amp, sigma, expect = [], [], [], [], [], []
    for s in range(S):
        text, bbbb = synth.random_message(rng, n_lines=2, words_per_line=4)
        b = synth.message_bits(text, n_phasing=30, n_tail=6)
        dur = len(b) / 100.0
        bits.append(b)
        ch = s % 2
        off.append(14000.0 if ch == 0 else -14000.0)
        start.append(0.3 + max(0.0, seconds - dur - 1.5) * rng.random())
        amp.append(3000.0 + 6000.0 * rng.random())
        sigma.append(amp[-1] * 10 ** (rng.uniform(-6.0, 14.0) / 20) / np.sqrt(2))
        if start[-1] + dur + 0.5 < seconds:
            expect.append((s + rank * S, 518 if ch == 0 else 490, bbbb, text))
    bufs = [torch.empty((S, block, 2), dtype=torch.float32, device=device) for _ in range(2)]
    eng = engine.Engine(S, block, device=local, first_stream_id=rank * S)
    es = torch.cuda.ExternalStream(eng.stream, device=device)

    def generate(k):
        engine.synth_fill_device(local, bufs[k % 2].data_ptr(), S, k * block, block, bits, off, start, amp, sigma,
                                 seed=65536 + rank, cuda_stream=eng.stream)

    generate(0)
    eng.push_device(bufs[0].data_ptr(), block)       # warm-up (engine reset afterwards)
    eng.poll_messages()
    eng.reset()
    eng.enable_timing(1)
    eng.stats()
    barrier()
    # The generator call is host-synchronous (it ends with a device-wide sync), so the engine's stream is idle when a
    # block is pushed: events on that stream around the push bracket exactly the block's cascade, tail carry and
    # feed-forward demod kernels; the block's sequential kernels run on the demod stream beside the next generator
    # call (in the resident benchmark: beside the next cascade).  The last block's remainder is added at the end.
    spans = []
    t_wall0 = time.time()
    for k in range(n_blocks):
        generate(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(es)
        eng.push_device(bufs[k % 2].data_ptr(), block)
        e1.record(es)
        spans.append((e0, e1))
    e2 = torch.cuda.Event(enable_timing=True)
    eng.sync()
    e2.record(es)
    torch.cuda.synchronize()
    t_all = (time.time() - t_wall0) * 1e3
    dec_ms = sum(a.elapsed_time(b) for a, b in spans) + spans[-1][1].elapsed_time(e2)
    t = torch.tensor([dec_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dec_ms = float(t.item())
    st = eng.stats()
    msgs = eng.poll_messages()
    got = {(m[0], m[1], m[2], m[3]) for m in msgs}
    ok = sum(1 for e in expect if e in got)
    merged = sharding.gather_messages(msgs)
    if rank != 0:
        return
    total = world * S * block * n_blocks
    peak, peak_src = measured_peak()
    casc_ms = st.cascade_ms / max(1, st.cascade_launches)
    print(json.dumps({
        "metric": "iq_msamples_per_s", "value": total / (dec_ms * 1e-3) / 1e6, "unit": "Msamples/s", "n_gpus": world,
        "steps": n_blocks, "warmup": 1, "ms_per_step": dec_ms / n_blocks, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[3]: 8192 streams per GPU x %.1f s of traffic, generated on the device per 1.25 s block" % seconds,
                   "streams_per_gpu": S, "samples_per_stream_per_step": block, "blocks": n_blocks,
                   "input": "float2 IQ, 20.6 GB per block per GPU, regenerated every block (nothing cache-warm)",
                   "parallelism": f"stream-sharded x{world}, no collectives"},
        "timing": {"decode_ms": dec_ms, "wall_ms_including_generation": t_all,
                   "how": "sum over blocks of the CUDA-event span of each push on the engine stream (generator untimed) + drain of the last block"},
        "realtime_streams": total / (dec_ms * 1e-3) / 252000,
        "roofline": {"bound": "hbm", "achieved": S * block * BYTES_PER_SAMPLE / (casc_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": S * block * BYTES_PER_SAMPLE / (casc_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "nvx::fir_cascade_kernel<true,false,false>", "kernel_ms": casc_ms},
        "gpu_launches": int(st.cascade_launches + st.demod_launches + st.aux_launches),
        "check": {"bulletins_complete_in_capture": len(expect), "decoded_exact": ok, "messages_total": len(msgs),
                  "messages_gathered_all_ranks": len(merged) if merged is not None else 0},
    }), flush=True)


def run_config5(args, torch, dist, device, rank, world, local, barrier):
    """BASELINE.json configs[4]: long-tap FIR stress.  The default workload's 1024 captures per GPU, resident in HBM, through
    --taps-tap Kaiser designs in all three stages (steeper than fir1cpp.C:8 / fir2cpp.C:22 / fir3cpp.h:16): stages 1 and 2 run
    on the tensor cores (fir_long_tc.cu), stage 3 and the demod chain as usual; every bulletin is checked."""
    from scipy import signal
    from navtex_b200 import engine, sharding

    S, T = STREAMS_PER_GPU, args.taps
    taps = (signal.firwin(T, 20000, window=("kaiser", 8.0), fs=252000), signal.firwin(T, 2000, window=("kaiser", 8.0), fs=63000),
            signal.firwin(T, 250, window=("kaiser", 7.0), fs=9000))
    x, expect = build_workload(torch, device, rank)
    eng = engine.Engine(S, BLOCK, device=local, first_stream_id=rank * S, taps=taps)
    es = torch.cuda.ExternalStream(eng.stream, device=device)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        eng.push_device(x.data_ptr(), BLOCK)
    eng.poll_messages()
    eng.reset()
    eng.enable_timing(1)
    eng.stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(es)
    for _ in range(args.steps):
        eng.push_device(x.data_ptr(), BLOCK)
    eng.sync()
    e1.record(es)
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    st = eng.stats()
    msgs = eng.poll_messages()          # the first pass decodes every bulletin; later passes continue the same streams
    got = {(m[0], m[1], m[2], m[3]) for m in msgs}
    ok = sum(1 for e in expect if e in got)
    merged = sharding.gather_messages(msgs)
    if rank != 0:
        return
    total = world * S * BLOCK * args.steps
    flop = 4.0 * T * (1 / 4 + 2 / 28 + 2 / 280) + 6 / 4          # per input sample: real FMAs x 2 per complex-by-real tap, + the mix
    fir_ms = st.cascade_ms / max(1, st.cascade_launches)         # the three stage kernels of one block (CUDA events)
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["bf16_tflops"]) / 2, "measured dense bf16 TF/s (MEASURED_PEAKS.json) / 2 = TF32 rate"
    except Exception:
        peak, peak_src = 1125.0, "fallback: nominal dense TF32 = 2250 / 2 TFLOP/s (B200_PROFILING.md)"
    ach = S * BLOCK * flop / (fir_ms * 1e-3) / 1e12

    def executed_tflop(D, rows, n_in, n_tile):
        # what the tensor cores execute for one stage (fir_long_tc.cu geometry): per 128-row tile of n_tile outputs, `chunks`
        # K chunks of 2 planes x 3 TF32 terms x 4 k-steps of M128 x N x K8
        t_pad = T
        while (D - t_pad) & 3:
            t_pad += 1
        cs = 32 // D * D
        chunks = -(-(D * (n_tile - 1) + t_pad) // cs)
        tiles = -(-rows // 128) * -(-(n_in // D) // n_tile)
        return tiles * chunks * 24 * 2.0 * 128 * n_tile * 8 / 1e12
    ex = executed_tflop(4, S, BLOCK, 128 if 384 <= T <= 548 else 64) + (executed_tflop(7, 2 * S, BLOCK // 4, 64) if T <= 959 else 0.0)
    print(json.dumps({
        "metric": "iq_msamples_per_s", "value": total / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (stages 1-2: 3xTF32 on tcgen05, FP32 accumulate)", "data": "synthetic",
        "config": {"workload": "configs[4]: long-tap FIR stress, %d-tap Kaiser designs in all three stages, 1024 synthetic IQ streams per GPU "
                               "resident in HBM (21.2 GB float2 per step: larger than L2)" % T,
                   "streams_per_gpu": S, "samples_per_stream_per_step": BLOCK, "taps": [T, T, T],
                   "parallelism": f"stream-sharded x{world}, no collectives"},
        "realtime_streams": total / (ms * 1e-3) / 252000,
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": "nvx::fir_tc_kernel<4,64> + <7,64> + fir_long_kernel<10,2>", "kernel_ms": fir_ms,
                     "note": "achieved = ALGORITHMIC FP32 flops (%.0f per input sample) / time of the three stage kernels; the tensor cores "
                             "execute 3x that for the TF32 split plus the structural zeros of the Toeplitz band" % flop,
                     "executed_tensor_tflops": ex / (fir_ms * 1e-3), "executed_frac": ex / (fir_ms * 1e-3) / peak},
        "e2e": None, "cpu_baseline": None,
        "gpu_launches": int(st.cascade_launches + st.demod_launches + st.aux_launches), "clocks": clocks,
        "check": {"bulletins_in_capture": len(expect), "decoded_exact": ok, "messages_total": len(msgs),
                  "messages_gathered_all_ranks": len(merged) if merged is not None else 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="config2", choices=["config2", "config4", "config5"],
                    help="config2 (default, the metric's configuration): 1024 streams/GPU resident; config4: 8192 streams/GPU x "
                         "--seconds of traffic generated on the device block by block (BASELINE.json configs[3]); config5: the "
                         "default captures through --taps-tap filters in all three stages (BASELINE.json configs[4])")
    ap.add_argument("--taps", type=int, default=255, help="config5: taps per stage")
    ap.add_argument("--seconds", type=float, default=60.0, help="config4: traffic per stream")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        impl_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from navtex_b200 import engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU path; use --impl reference for the CPU chain)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner (and anything else NCCL_DEBUG asks for) to
        # stdout unless told otherwise
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload in ("config4", "config5"):
        (run_config4 if args.workload == "config4" else run_config5)(args, torch, dist, device, rank, world, local, barrier)
        if world > 1:
            dist.destroy_process_group()
        return

    S = STREAMS_PER_GPU
    x, expect = build_workload(torch, device, rank)
    eng = engine.Engine(S, BLOCK, device=local, first_stream_id=rank * S)
    es = torch.cuda.ExternalStream(eng.stream, device=device)

    # ---- device-resident whole-job throughput ----------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        eng.push_device(x.data_ptr(), BLOCK)
    msgs_warm = eng.poll_messages()
    eng.enable_timing(1)            # two event records per block around the fused FIR kernel, nothing else
    eng.stats()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record(es)
    for _ in range(args.steps):
        eng.push_device(x.data_ptr(), BLOCK)
    eng.sync()
    ev1.record(es)
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    dev_ms = ev0.elapsed_time(ev1)
    st = eng.stats()
    msgs = eng.poll_messages()
    # demod stage breakdown from a short extra pass (the per-stage event records would only add bubbles to the timed region)
    eng.enable_timing(2)
    eng.stats()
    for _ in range(3):
        eng.push_device(x.data_ptr(), BLOCK)
    st2 = eng.stats()
    eng.poll_messages()
    eng.enable_timing(0)
    # correctness of what was timed: each pass over the block re-decodes every stream's bulletin
    got = {(m[0], m[1], m[2], m[3]) for m in msgs}
    decoded_ok = sum(1 for e in expect if e in got)
    # the only cross-rank data exchange of the job: final host gather of the decoded message records (untimed)
    from navtex_b200 import sharding
    merged = sharding.gather_messages(msgs)
    gathered = len(merged) if merged is not None else 0

    t = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    total_samples = world * S * BLOCK * args.steps
    value = total_samples / (max_ms * 1e-3) / 1e6                     # Msamples/s

    # ---- secondary: the same captures resident as int16 I,Q (the radio's own format), fused-ingest kernel variant ----
    x16 = x.round().to(torch.int16)
    eng16 = engine.Engine(S, BLOCK, device=local, first_stream_id=rank * S)
    s16s = torch.cuda.ExternalStream(eng16.stream, device=device)
    for _ in range(args.warmup):
        eng16.push_device(x16.data_ptr(), BLOCK, s16=True)
    eng16.poll_messages()
    eng16.enable_timing(1)
    eng16.stats()
    k16 = min(args.steps, 20)
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(s16s)
    for _ in range(k16):
        eng16.push_device(x16.data_ptr(), BLOCK, s16=True)
    eng16.sync()
    b1.record(s16s)
    barrier()
    st16 = eng16.stats()
    got16 = {(m[0], m[1], m[2], m[3]) for m in eng16.poll_messages()}
    t16 = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t16, op=dist.ReduceOp.MAX)
    casc16_ms = st16.cascade_ms / max(1, st16.cascade_launches)
    int16_input = {
        "value": world * S * BLOCK * k16 / (float(t16.item()) * 1e-3) / 1e6, "unit": "Msamples/s", "steps": k16,
        "ms_per_step": float(t16.item()) / k16, "kernel": "nvx::fir_cascade_kernel<true,false,true> (short2 rows by TMA, PRMT/FADD2 conversion)",
        "kernel_ms": casc16_ms, "algorithmic_bytes_per_sample": 4.0 + 16.0 / 280,
        "achieved_gbs": S * BLOCK * (4.0 + 16.0 / 280) / (casc16_ms * 1e-3) / 1e9, "bound": "fp32 issue (not HBM)",
        # 55.5 algorithmic flop per input sample against 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.5 TFLOP/s
        "fp32_tflops": S * BLOCK * 55.5 / (casc16_ms * 1e-3) / 1e12, "fp32_frac_of_peak": S * BLOCK * 55.5 / (casc16_ms * 1e-3) / 74.5e12,
        "decoded_exact": sum(1 for e in expect if e in got16),
    }
    eng16.close()
    del x16, eng16

    cpu_sample = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = 4 * 252000
        cpu_sample = [x[k, :n_cpu].round().to(torch.int16).cpu().numpy().reshape(-1) for k in range(min(os.cpu_count() or 1, 16))]

    # ---- end to end through the host-buffer C ABI --------------------------------------------------
    # The same 1024 captures, as the reference's int16 samples in pinned host memory, pushed as consecutive 1.03 s
    # blocks (step k pushes chunk k mod 10; after the tenth the captures start over, like a new emission).
    ne = E2E_BLOCK
    assert ne % 280 == 0 and ne * E2E_CHUNKS == BLOCK
    host = torch.empty((E2E_CHUNKS, S, ne, 2), dtype=torch.int16).pin_memory()
    for k in range(E2E_CHUNKS):
        host[k].copy_(x[:, k * ne:(k + 1) * ne].round().to(torch.int16))
    del x
    torch.cuda.empty_cache()
    e2e_eng = engine.Engine(S, ne, device=local, first_stream_id=rank * S)
    e2s = torch.cuda.ExternalStream(e2e_eng.stream, device=device)
    warm = E2E_CHUNKS * max(1, (max(3, args.warmup) + E2E_CHUNKS - 1) // E2E_CHUNKS)     # whole captures, so step 0 starts one
    for k in range(warm):
        e2e_eng.push_host_ptr(host[k % E2E_CHUNKS].data_ptr(), ne, s16=True)
        e2e_eng.poll_messages()
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    a0.record(e2s)
    e2e_msgs = []
    for k in range(args.steps):
        e2e_eng.push_host_ptr(host[k % E2E_CHUNKS].data_ptr(), ne, s16=True)
        # every block's events are downloaded (D2H) and assembled on the host as part of its push; collect what has
        # completed so far without stalling the copy / compute pipeline
        e2e_msgs += e2e_eng.poll_messages(wait=False)
    e2e_msgs += e2e_eng.poll_messages()                               # drain: the last blocks' results, inside the timed region
    a1.record(e2s)
    barrier()
    e2e_wall = time.time() - tw0
    te = torch.tensor([max(a0.elapsed_time(a1) * 1e-3, e2e_wall)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * S * ne * args.steps / float(te.item()) / 1e6
    ev_cap = 2 * (ne // 280 // 63 + 2) + 8
    e2e_expected = len(expect) * (args.steps // E2E_CHUNKS)
    expect_set = set(expect)
    e2e_exact = sum(1 for m in e2e_msgs if (m[0], m[1], m[2], m[3]) in expect_set)
    e2e = {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": S * ne * 4, "d2h_bytes_per_step": 2 * S * (ev_cap + 4),
           "input": f"int16 IQ in pinned host memory, [{S} streams][{ne} samples] per step, consecutive blocks of the same captures",
           "ms_per_step": float(te.item()) * 1e3 / args.steps,
           "numa_binding": ("node %d, %d cpus" % numa) if numa else None}
    e2e_eng.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused FIR cascade) -------------------------------------------
    peak, peak_src = measured_peak()
    casc_ms = st.cascade_ms / max(1, st.cascade_launches)
    achieved = S * BLOCK * BYTES_PER_SAMPLE / (casc_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic["dram_bytes_per_launch"] if traffic else None, "peak_source": peak_src,
                "kernel": "nvx::fir_cascade_kernel<true,false,false>", "kernel_ms": casc_ms,
                "demod_chain_ms": st2.demod_ms / max(1, st2.cascade_launches),
                "demod_stage_ms": dict(zip(("angle_corr", "offset_sum", "carry", "symbol_clock", "bit_decide", "fsm"),
                                           (v / max(1, st2.cascade_launches) for v in st2.demod_stage_ms))),
                "algorithmic_bytes_per_launch": S * BLOCK * BYTES_PER_SAMPLE,
                "kernel_gsamples_per_s": S * BLOCK / (casc_ms * 1e-3) / 1e9}

    # ---- CPU baseline: the reference's own chain on the host cores, bounded sample ------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = 4 * 252000
        sample = cpu_sample
        if os.path.exists(REF_CHAIN):
            sps, step_s, n = run_reference_cpu(sample, 2, 10, cores)
            kind = "reference"
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as ol
            t0 = time.time()
            for k in range(4):
                ol.run_oracle(sample[k % len(sample)], record_taps=False)
            sps, cores, kind = 4 * n_cpu / (time.time() - t0), 1, "port"
        cpu = {"value": sps / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
               "sample": f"first 4 s of {min(cores, 16)} of the workload's streams, one ref_chain process per core, 10 timed passes"}

    line = {
        "metric": "iq_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "configs[1]: 1024 synthetic IQ streams per GPU through fused FIR cascade + FSK demod + bit-sync + SITOR-B",
                   "streams_per_gpu": S, "samples_per_stream_per_step": BLOCK, "seconds_per_step": BLOCK / 252000,
                   "input": "float2 IQ resident in HBM, 21.2 GB per GPU per step (larger than L2; no flush needed)",
                   "parallelism": f"stream-sharded x{world}, no collectives"},
        "realtime_streams": value / 0.252,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "int16_input": int16_input,
        "gpu_launches": int(st.cascade_launches + st.demod_launches + st.aux_launches),
        "clocks": clocks,
        "check": {"bulletins_expected_per_step": len(expect), "decoded_exact": decoded_ok, "messages_total": len(msgs), "messages_gathered_all_ranks": gathered,
                  "e2e_messages": len(e2e_msgs), "e2e_messages_exact": e2e_exact, "e2e_bulletins_completed_in_timed_steps": e2e_expected},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
