#!/bin/sh
# Round-2 closing evidence on one GPU: GPU tests, smoke, the default bench line, the ncu passes of tools/run_profile_r2.sh, and the
# configs[4] bench lines + launch list with the stage-2 mix on load.
set -x
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1; tail -3 gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 600 gpurun_out/final_bench_n1.json
timeout 1200 sh tools/run_profile_r2.sh > gpurun_out/final_profile.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_config5.csv python bench.py --workload config5 --taps 255 --steps 2 --warmup 1 > gpurun_out/ncu_c5.log 2>&1
for t in 65 127 255 383 511; do timeout 200 python bench.py --workload config5 --taps $t --steps 10 --warmup 3 > gpurun_out/final_c5_$t.json 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep gpurun_out/final_*
