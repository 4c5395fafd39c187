python bench.py --workload config5 --taps ${T:-65} --steps 1 --warmup 1 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fir_long_kernel -s 1 -c 1 -o gpurun_out/r2_stage3 python bench.py --workload config5 --taps ${T:-65} --steps 1 --warmup 1 > gpurun_out/r2_ncu_stage3.log 2>&1
