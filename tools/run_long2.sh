# development aid: load-policy sweep of the tensor-core long-tap kernel + per-kernel launch times
for ld in 0 1 2; do
  echo "== taps 255 N=128 NVX_TC_LD=$ld"
  NVX_LONG_TC=1 NVX_TC_LD=$ld timeout 120 python tools/quick_perf.py --steps 5 --timing 1 --taps 255 --super 4625 2>&1 | tail -1
done
for m in 0 1; do
NVX_LONG_TC=$m ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fir_" -s 6 -c 3 --csv --log-file gpurun_out/launch_tc$m.csv \
   python tools/quick_perf.py --steps 1 --taps 255 --super 4625 > /dev/null 2>&1
grep fir_ gpurun_out/launch_tc$m.csv | awk -F'","' '{print $5, $NF}' | cut -c1-150
done
