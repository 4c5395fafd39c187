free -g | head -2; nproc
for mode in default overlap default overlap; do
  if [ $mode = default ]; then unset NVX_PIPELINE; else export NVX_PIPELINE=$mode; fi
  echo "== pipeline=$mode"
  python tools/quick_perf.py --steps 40 --timing 1 2>&1 | tail -2 | grep -v stages
done
