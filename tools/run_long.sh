# development aid: long-tap parity tests and the tap-count sweep behind DESIGN.md 3.4
python -m pytest tests/test_gpu_longtaps.py -m gpu -x -q 2>&1 | tail -2
for t in 63 127 255 511; do
  echo "== taps $t"
  python tools/quick_perf.py --steps 5 --timing 1 --taps $t --super 4625 2>&1 | tail -3 | grep -v stages
done
