for fmt in "" "--s16"; do
echo "== NVX_LONG_TC=1 $fmt"
NVX_LONG_TC=1 timeout -s KILL 120 python tools/quick_perf.py --steps 5 --timing 1 --taps 255 --super 4625 $fmt 2>&1 | tail -1
NVX_LONG_TC=1 timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fir_" -s 6 -c 3 --csv --log-file gpurun_out/launch_tc1.csv \
   python tools/quick_perf.py --steps 1 --taps 255 --super 4625 $fmt > /dev/null 2>&1
grep fir_ gpurun_out/launch_tc1.csv | awk -F'","' '{print $5, $NF}' | cut -c1-150
done
