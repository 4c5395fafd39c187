for mode in default ffmain default ffmain; do
  if [ $mode = default ]; then unset NVX_PIPELINE; else export NVX_PIPELINE=$mode; fi
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv,noheader,nounits -lms 20 > /tmp/clk_$mode.csv &
  SMI=$!
  python tools/quick_perf.py --steps 400 --timing 1 2>&1 | tail -1
  kill $SMI
  echo "== pipeline=$mode clocks (last 60 samples = under load):"
  tail -60 /tmp/clk_$mode.csv | awk -F, '{c[$1]++; p+=$2; n++} END {for (k in c) printf "%s MHz x%d  ", k, c[k]; printf " avg power %.0f W\n", p/n}'
done
