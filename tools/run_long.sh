# development aid: tap-length parity tests and the tap-count sweep behind DESIGN.md 3.4 (fused medium class <= 61, long path beyond)
timeout -s KILL 300 python -m pytest tests/test_gpu_longtaps.py -m gpu -q -x 2>&1 | tail -4
for t in 127 255 511; do
  for m in "1 128" "1 64" "0 0"; do
    set -- $m
    echo "== taps $t NVX_LONG_TC=$1 NVX_TC_N=$2"
    NVX_LONG_TC=$1 NVX_TC_N=$2 timeout -s KILL 120 python tools/quick_perf.py --steps 5 --timing 1 --taps $t --super 4625 2>&1 | tail -3 | grep -v stages
  done
done
