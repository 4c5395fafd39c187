for d in 0 32; do
  echo "== dbg $d"
  NVX_LONG_TC=1 NVX_TC_DBG=$d timeout -s KILL 60 python tools/quick_perf.py --steps 1 --timing 1 --taps 255 --super 500 2>&1 | tail -1 | cut -c1-200
done
