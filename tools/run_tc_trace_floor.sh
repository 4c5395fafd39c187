# issuer timeline of the streaming stage-2 kernel with everything else knocked out (NVX_TC_DBG=15) and in the full pipeline
for dbg in 15 0; do
echo "== dbg $dbg"
NVX_TC_TRACE=${D:-7} NVX_TC_DBG=$dbg NVX_TC_TRACE_FULL=1 python bench.py --workload config5 --taps ${T:-65} --steps 3 --warmup 1 2>&1 >/dev/null | grep -v Traceback | sed -n '1,2p;30,52p'
done
