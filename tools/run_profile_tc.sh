set -x
# ncu capture of the tensor-core long-tap kernels (255 taps, 1024 streams x 1,295,000 samples = one 5.1 s block): launch list of
# the three stages, then --set full of fir_tc_kernel<4, 64> (stage 1) and fir_tc_kernel<7, 64> (stage 2)
python tools/quick_perf.py --steps 2 --taps ${TAPS:-255} --super 4625 > gpurun_out/plain_tc.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fir_" -s 6 -c 3 --csv --log-file gpurun_out/launch_tc.csv \
    python tools/quick_perf.py --steps 1 --taps ${TAPS:-255} --super 4625 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fir_tc" -s 4 -c 2 \
    -o gpurun_out/prof_tc python tools/quick_perf.py --steps 1 --taps ${TAPS:-255} --super 4625 > gpurun_out/ncu_tc.log 2>&1
ls -la gpurun_out/prof_tc.ncu-rep
