# A/B of "mix on load" (stage 2 rotates while it converts; default) against the rotation in the stage-1 epilogue (NVX_LONG_MIX=stage1)
set -x
python -m pytest tests/test_gpu_longtaps.py tests/test_long_tc_band.py -x -q -m gpu 2>&1 | tail -5
for t in ${TAPS:-65 255}; do
  python bench.py --workload config5 --taps $t --steps 10 --warmup 3 > gpurun_out/mix_new_$t.json 2> gpurun_out/mix_new_$t.err
  NVX_LONG_MIX=stage1 python bench.py --workload config5 --taps $t --steps 10 --warmup 3 > gpurun_out/mix_old_$t.json 2> gpurun_out/mix_old_$t.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 -k regex:fir_ --csv --log-file gpurun_out/mix_launches.csv python bench.py --workload config5 --taps 255 --steps 1 --warmup 1 > gpurun_out/ncu_c5.log 2>&1
for f in gpurun_out/mix_*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value']/1e3,1), 'Gs/s', d['ms_per_step'], d['roofline']['kernel_ms'], d['check'], d['clocks']['sm_mhz'])
PY
done
grep -h "fir_" gpurun_out/mix_launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | head -20
