# development aid: GPU parity tests, then tools/quick_perf.py in a few configurations
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for args in "" "--s16" "--nco" "--s16 --nco"; do
  echo "== quick_perf $args"
  python tools/quick_perf.py --steps 20 --timing 1 $args 2>&1 | tail -3
done
