export NVX_PIPELINE=overlap
for pad in 0 16000 24000 38000; do
  echo "== overlap pad=$pad"
  NVX_DEMOD_PAD=$pad python tools/quick_perf.py --steps 40 --timing 2 2>&1 | tail -3 | cut -c1-140
done
unset NVX_PIPELINE
echo "== default"; python tools/quick_perf.py --steps 40 --timing 2 2>&1 | tail -3 | cut -c1-140
