# development aid: GPU parity tests, then the pipeline-mode matrix of tools/quick_perf.py
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for mode in default overlap; do
  if [ $mode = default ]; then unset NVX_PIPELINE; else export NVX_PIPELINE=$mode; fi
  echo "== pipeline=$mode"
  python tools/quick_perf.py --steps 10 2>&1 | tail -3
done
