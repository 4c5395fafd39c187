for m in 3 1; do
echo "== taps 1023 NVX_LONG_TC=$m"
NVX_LONG_TC=$m timeout -s KILL 120 python tools/quick_perf.py --steps 3 --timing 1 --taps 1023 --super 4625 2>&1 | tail -1
done
for m in 3 1; do
echo "== taps 767 NVX_LONG_TC=$m"
NVX_LONG_TC=$m timeout -s KILL 120 python tools/quick_perf.py --steps 3 --timing 1 --taps 767 --super 4625 2>&1 | tail -1
done
