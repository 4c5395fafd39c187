for d in 0 1 2 8 16 24 27; do
echo "== dbg $d"
NVX_TC_DBG=$d timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"fir_tc" -s 2 -c 1 --csv --log-file gpurun_out/launch_dbg.csv \
   python tools/quick_perf.py --steps 1 --taps 255 --super 4625 > /dev/null 2>&1
grep fir_ gpurun_out/launch_dbg.csv | awk -F'","' '{print $(NF-2), $NF}' | cut -c1-150
done
