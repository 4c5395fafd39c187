for r in 4 12 20 4 12 20; do
  echo "== reserve=$r"
  NVX_RESERVE_SMS=$r python tools/quick_perf.py --steps 40 --timing 1 2>&1 | tail -1
done
