for v in "" s16w8 s16w8s1 s16w8roll roll; do
  if [ -z "$v" ]; then unset NVX_LIB; else export NVX_LIB=$PWD/navtex_b200/variants/libnavtex_b200_$v.so; fi
  echo "== variant '$v' s16"
  python tools/quick_perf.py --steps 20 --timing 1 --s16 2>&1 | tail -3 | grep -v stages
done
export NVX_LIB=$PWD/navtex_b200/variants/libnavtex_b200_roll.so
echo "== variant roll f32"
python tools/quick_perf.py --steps 20 --timing 1 2>&1 | tail -3 | grep -v stages
