timeout 300 python -m pytest tests/test_gpu_longtaps.py -x -q -m gpu 2>&1 | tail -2
for t in 65 255; do timeout 200 python bench.py --workload config5 --taps $t --steps 10 --warmup 3 > gpurun_out/s3_$t.json 2>gpurun_out/s3_$t.err; tail -3 gpurun_out/s3_$t.err; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 -k regex:fir_long_kernel --csv --log-file gpurun_out/s3_launches.csv python bench.py --workload config5 --taps 65 --steps 1 --warmup 1 > /dev/null 2>&1
grep fir_long gpurun_out/s3_launches.csv | awk -F'","' '{print $NF}' | head -4
for f in gpurun_out/s3_*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']
print(sys.argv[1], round(d['value']/1e3,1), 'Gs/s', round(d['ms_per_step'],3), round(r['kernel_ms'],3), d['check']['decoded_exact_all_ranks'], d['clocks']['sm_mhz'], 'exec_frac', round(r['executed_frac'],3), 'hbm', r['hbm']['y1_rows_per_stream'], round(r['hbm']['algorithmic_bytes_per_sample'],3), round(r['hbm']['frac'],3))
PY
done
