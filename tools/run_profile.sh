# development aid: the ncu passes behind profiles/ (launch list of the bench command, full captures of the hot kernels)
set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_r1d.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r1d.log 2>&1
python tools/quick_perf.py --steps 1 > gpurun_out/plain_qp.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fir_cascade|offset_sum|angle_corr|symbol_clock|fsm_kernel|bit_decide" -s 12 -c 6 \
    -o gpurun_out/prof_r1d python tools/quick_perf.py --steps 1 > gpurun_out/ncu_qp.log 2>&1
python tools/quick_perf.py --steps 1 --s16 > gpurun_out/plain_qp16.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fir_cascade" -s 2 -c 1 \
    -o gpurun_out/prof_r1d_s16 python tools/quick_perf.py --steps 1 --s16 > gpurun_out/ncu_qp16.log 2>&1
python tools/quick_perf.py --steps 1 --taps 255 --super 2000 > gpurun_out/plain_qpl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fir_long" -s 6 -c 3 \
    -o gpurun_out/prof_r1d_long python tools/quick_perf.py --steps 1 --taps 255 --super 2000 > gpurun_out/ncu_qpl.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/ncu_r1d.log gpurun_out/ncu_qp.log gpurun_out/ncu_qp16.log gpurun_out/ncu_qpl.log
