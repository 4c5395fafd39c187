for v in "" s16w12 s16w16; do
  if [ -z "$v" ]; then unset NVX_LIB; else export NVX_LIB=$PWD/navtex_b200/variants/libnavtex_b200_$v.so; fi
  echo "== variant '$v' s16"
  python tools/quick_perf.py --steps 20 --timing 1 --s16 2>&1 | tail -1
done
