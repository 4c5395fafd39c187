# development aid: tap-length parity tests and the tap-count sweep behind DESIGN.md 3.4 (fused medium class <= 61, long path beyond)
timeout -s KILL 400 python -m pytest tests/test_gpu_longtaps.py -m gpu -q -x 2>&1 | tail -4
for t in ${TAPS:-65 95 127 255 511 1023}; do
  for m in ${MODES:-1 0}; do
    for fmt in "" "--s16"; do
    echo "== taps $t NVX_LONG_TC=$m $fmt"
    NVX_LONG_TC=$m timeout -s KILL 120 python tools/quick_perf.py --steps 5 --timing 1 --taps $t --super 4625 $fmt 2>&1 | tail -1
    done
  done
done
