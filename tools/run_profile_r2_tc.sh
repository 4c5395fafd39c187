set -x
python bench.py --workload config5 --taps 255 --steps 1 --warmup 1 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fir_tcs_kernel -s 2 -c 2 -o gpurun_out/r2_tcs python bench.py --workload config5 --taps 255 --steps 1 --warmup 1 > gpurun_out/r2_ncu_tcs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_config5.csv python bench.py --workload config5 --taps 255 --steps 2 --warmup 1 > gpurun_out/ncu_c5.log 2>&1
for t in 65 127 255 511; do python bench.py --workload config5 --taps $t --steps 10 --warmup 2 > gpurun_out/r2_c5_final_$t.json 2>/dev/null; done
