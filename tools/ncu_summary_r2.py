#!/usr/bin/env python3
"""Turn the round-2 ncu captures (tools/run_profile_r2.sh, brought back in gpurun_out/) into the tracked evidence under profiles/:
raw CSV pages, a launch-list digest, profiles/cascade_traffic.json (DRAM bytes per launch of the fused kernel, with the command,
commit and date it was measured at -- bench.py copies those fields into roofline.traffic_source) and profiles/r2_ncu_summary.md."""
import collections
import csv
import datetime
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]


def raw_page(rep, dst):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(dst, "w").write(txt)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for k in WANT:
            if k in hdr:
                d[k] = (r[hdr.index(k)], units[hdr.index(k)])
        out.append(d)
    return out


def num(x):
    return float(x.replace(",", ""))


def to_bytes(v, unit):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def to_ms(v, unit):
    return num(v) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(unit, 1.0))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(to_ms(r[vi], r[ui]))
    return agg


def main():
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip()
    date = datetime.datetime.utcnow().strftime("%Y-%m-%d")
    md = ["# Round-2 ncu evidence (tools/run_profile_r2.sh; commit %s, %s)\n" % (commit, date),
          "Numbers under ncu are cold-cache and serialised: shares and ratios count, absolute times are the bench's (CUDA events).\n"]
    # ---- launch list
    ll = os.path.join(OUT, "r2_launches_bench.csv")
    if os.path.exists(ll):
        subprocess.run(["cp", ll, os.path.join(PROF, "r2_launches_bench.csv")])
        agg = launches(ll)
        md.append("## Launch list of `python bench.py --steps 3 --warmup 1` (profiles/r2_launches_bench.csv)\n")
        md.append("| kernel | launches | mean ms | min ms |\n|---|---|---|---|")
        for k, v in agg.items():
            md.append("| `%s` | %d | %.3f | %.3f |" % (k[:90], len(v), sum(v) / len(v), min(v)))
        md.append("")
    # ---- cascade
    rep = os.path.join(OUT, "r2_cascade.ncu-rep")
    if os.path.exists(rep):
        rows = raw_page(rep, os.path.join(PROF, "r2_ncu_full_raw_cascade.csv"))
        md.append("## Fused FIR kernel on the bench workload (profiles/r2_ncu_full_raw_cascade.csv)\n")
        for d in rows:
            md.append("* `%s`: " % d["Kernel Name"][:80] + "; ".join("%s = %s %s" % (k, d[k][0], d[k][1]) for k in WANT if k in d))
        d = rows[-1]
        rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
        alg = 1024 * 2590000 * (8.0 + 16.0 / 280)
        traffic = {"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "algorithmic_bytes_per_launch": alg,
                   "ratio_to_algorithmic": (rd + wr) / alg, "kernel": d["Kernel Name"],
                   "source": "profiles/r2_ncu_full_raw_cascade.csv (ncu --set full --clock-control none, second timed launch)",
                   "command": "python bench.py --steps 2 --warmup 1 --skip parity,int16,e2e,config4,config5,channels,cpu  (the bench workload: 1024 streams x 2,590,000 samples)",
                   "commit": commit, "date": date}
        json.dump(traffic, open(os.path.join(PROF, "cascade_traffic.json"), "w"), indent=1)
        md.append("\nDRAM traffic per launch %.3f GB = %.3f x the algorithmic 21.369 GB -> profiles/cascade_traffic.json\n" % ((rd + wr) / 1e9, (rd + wr) / alg))
    # ---- tensor-core stages
    rep = os.path.join(OUT, "r2_tcs.ncu-rep")
    if os.path.exists(rep):
        rows = raw_page(rep, os.path.join(PROF, "r2_ncu_full_raw_tcs.csv"))
        md.append("## Streaming tensor-core long-tap stages, 255 taps (profiles/r2_ncu_full_raw_tcs.csv)\n")
        for d in rows:
            md.append("* `%s`: " % d["Kernel Name"][:80] + "; ".join("%s = %s %s" % (k, d[k][0], d[k][1]) for k in WANT if k in d))
        md.append("")
    open(os.path.join(PROF, "r2_ncu_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    sys.exit(main())
