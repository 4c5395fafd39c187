#!/bin/sh
# Round-2 ncu evidence, one GPU (run under gpurun; every command first runs WITHOUT ncu and must exit 0):
#   1. launch list of the default bench command (per-launch gpu__time_duration: the kernels' SHARE of a step)
#   2. ncu --set full of the fused FIR kernel on the bench workload  -> profiles/cascade_traffic.json (tools/ncu_summary_r2.py)
#   3. ncu --set full of the streaming tensor-core stages (config5, 255 taps)
set -x
SKIP=parity,int16,e2e,config4,config5,channels,cpu
python bench.py --steps 3 --warmup 1 > gpurun_out/r2_plain_bench.json 2> gpurun_out/r2_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 3 --warmup 1 > gpurun_out/r2_ncu_launches.log 2>&1
python bench.py --steps 2 --warmup 1 --skip $SKIP > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fir_cascade_kernel -s 1 -c 2 -o gpurun_out/r2_cascade \
    python bench.py --steps 2 --warmup 1 --skip $SKIP > gpurun_out/r2_ncu_cascade.log 2>&1
python bench.py --workload config5 --taps 255 --steps 1 --warmup 1 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fir_tcs_kernel -s 2 -c 2 -o gpurun_out/r2_tcs \
    python bench.py --workload config5 --taps 255 --steps 1 --warmup 1 > gpurun_out/r2_ncu_tcs.log 2>&1
ls -la gpurun_out/*.ncu-rep
