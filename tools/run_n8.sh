nvidia-smi topo -m 2>&1 | head -14
lscpu | grep -i "numa\|socket\|model name" | head -8
cat /sys/devices/system/node/node*/cpulist 2>/dev/null | head -4
python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8b.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'])"
