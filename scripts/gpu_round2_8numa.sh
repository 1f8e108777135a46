#!/bin/bash
# 8 GPUs: does binding every rank to its GPU's socket close the gap between the resident and the end-to-end number?
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02y_topo.txt 2>&1
python -c "import os; print('allowed cpus', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], '...')" >> gpurun_out/r02y_topo.txt
lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r02y_topo.txt 2>&1
for bind in 0 1; do
  B2W_NUMA_BIND=$bind python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus 8 --steps 3 --warmup 3 --no-workloads --no-cpu-baseline --parity-utts 0 > gpurun_out/r02y_bench_8gpu_bind$bind.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/r02y_bench_8gpu_bind$bind.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('bind', $bind, d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e_components'], d['config'].get('host_cpus_bound_to_gpu_socket'))
else:
    print(open('gpurun_out/r02y_bench_8gpu_bind$bind.log').read()[-2000:])
PY
done
tail -25 gpurun_out/r02y_topo.txt
