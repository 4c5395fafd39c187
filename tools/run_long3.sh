for d in 0 64; do
echo "== NVX_LONG_TC=1 dbg=$d"
NVX_LONG_TC=1 NVX_TC_DBG=$d timeout -s KILL 120 python tools/quick_perf.py --steps 5 --timing 1 --taps 255 --super 4625 2>&1 | tail -1
NVX_LONG_TC=1 NVX_TC_DBG=$d timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"fir_tc" -s 2 -c 1 --csv --log-file gpurun_out/launch_tc1.csv \
   python tools/quick_perf.py --steps 1 --taps 255 --super 4625 > /dev/null 2>&1
grep fir_ gpurun_out/launch_tc1.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-150
done
