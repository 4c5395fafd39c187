set -x
# ncu capture of the tensor-core long-tap stage-1 kernel (255 taps, 1024 streams x 1,295,000 samples = one 5.1 s block)
python tools/quick_perf.py --steps 2 --taps ${TAPS:-255} --super 4625 > gpurun_out/plain_tc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fir_tc" -s 2 -c 1 \
    -o gpurun_out/prof_tc python tools/quick_perf.py --steps 1 --taps ${TAPS:-255} --super 4625 > gpurun_out/ncu_tc.log 2>&1
ls -la gpurun_out/prof_tc.ncu-rep
