# development aid: GPU parity tests, then tools/quick_perf.py in the pipeline modes
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for mode in default overlap; do
  if [ $mode = default ]; then unset NVX_PIPELINE; else export NVX_PIPELINE=$mode; fi
  echo "== pipeline=$mode"
  python tools/quick_perf.py --steps 30 --timing 2 2>&1 | tail -3
done
