# in-kernel timeline of the streaming tensor-core kernel of one stage (NVX_TC_TRACE=<decimation>) with parts knocked out
# (NVX_TC_DBG bit mask: 1 no loads, 2 no conversion, 4 no tcgen05.st, 8 no epilogue; results are garbage then)
D=${D:-7}; T=${T:-255}
for dbg in 0 1 2 4 8 15; do
  echo "== dbg $dbg"
  NVX_TC_TRACE=$D NVX_TC_DBG=$dbg python bench.py --workload config5 --taps $T --steps 3 --warmup 1 2>&1 >/dev/null | grep "^#" | tail -2
done
NVX_TC_TRACE=$D NVX_TC_TRACE_FULL=1 python bench.py --workload config5 --taps $T --steps 3 --warmup 1 2> gpurun_out/tc_trace_D${D}_T${T}.txt >/dev/null
tail -n +1 gpurun_out/tc_trace_D${D}_T${T}.txt | grep -v "^Traceback" | sed -n 1,80p
