# development aid: ncu full captures of the kernels that changed after capture d (int16 cascade with 12 warps, long-tap kernels with cp.async staging)
set -x
python tools/quick_perf.py --steps 1 --s16 > gpurun_out/plain_qp16e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fir_cascade" -s 2 -c 1 \
    -o gpurun_out/prof_r1e_s16 python tools/quick_perf.py --steps 1 --s16 > gpurun_out/ncu_qp16e.log 2>&1
python tools/quick_perf.py --steps 1 --taps 255 --super 2000 > gpurun_out/plain_qple.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fir_long" -s 6 -c 3 \
    -o gpurun_out/prof_r1e_long python tools/quick_perf.py --steps 1 --taps 255 --super 2000 > gpurun_out/ncu_qple.log 2>&1
ls -la gpurun_out/prof_r1e*.ncu-rep
