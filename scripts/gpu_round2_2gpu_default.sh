#!/bin/bash
# the bench under torchrun with its DEFAULT workloads (rank 0 alone runs them: nothing in them may enter a collective)
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus 2 --utts 1024 --steps 2 --warmup 1 --no-cpu-baseline --io-utts 256 --parity-utts 2 > gpurun_out/r03f_bench_2gpu_workloads.log 2>&1
echo "exit $?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r03f_bench_2gpu_workloads.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['n_gpus'], d['value'], d['e2e']['value'], {k:(v.get('value') or v.get('failed') or 'ok') for k,v in d['workloads'].items()})
else:
    print(open('gpurun_out/r03f_bench_2gpu_workloads.log').read()[-2500:])
PY
