# development aid: per-kernel launch times of the long-tap path (255 taps), CUDA-core vs tensor-core stage 1, float2 and short2 input
for fmt in "" "--s16"; do
for m in 0 1; do
echo "== NVX_LONG_TC=$m $fmt"
NVX_LONG_TC=$m timeout -s KILL 120 python tools/quick_perf.py --steps 5 --timing 1 --taps 255 --super 4625 $fmt 2>&1 | tail -1
NVX_LONG_TC=$m timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fir_" -s 6 -c 3 --csv --log-file gpurun_out/launch_tc$m.csv \
   python tools/quick_perf.py --steps 1 --taps 255 --super 4625 $fmt > /dev/null 2>&1
grep fir_ gpurun_out/launch_tc$m.csv | awk -F'","' '{print $5, $NF}' | cut -c1-150
done
done
