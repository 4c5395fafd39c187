# development aid: tap-length parity tests and the tap-count sweep behind DESIGN.md 3.4 (fused medium class <= 61, long path beyond)
python -m pytest tests/test_gpu_longtaps.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python tools/sanitize_case.py 2>&1 | tail -3
for t in 37 47 61 63 127 255; do
  echo "== taps $t"
  python tools/quick_perf.py --steps 5 --timing 1 --taps $t --super 4625 2>&1 | tail -3 | grep -v stages
done
