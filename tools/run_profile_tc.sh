set -x
export NVX_LONG_TC=${TC_MASK:-1}
python tools/quick_perf.py --steps 1 --taps 255 --super 1000 > gpurun_out/plain_tc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fir_tc" -s 2 -c 1 \
    -o gpurun_out/prof_tc python tools/quick_perf.py --steps 1 --taps 255 --super 1000 > gpurun_out/ncu_tc.log 2>&1
ls -la gpurun_out/prof_tc.ncu-rep
