set -x
python bench.py --workload config5 --taps 255 --steps 1 --warmup 1 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fir_tcs_kernel -s 2 -c 2 -o gpurun_out/r2_tcs_mix python bench.py --workload config5 --taps 255 --steps 1 --warmup 1 > gpurun_out/r2_ncu_tcs_mix.log 2>&1
for t in 65 255; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_config5_mix_$t.csv python bench.py --workload config5 --taps $t --steps 2 --warmup 1 > gpurun_out/ncu_c5.log 2>&1
done
