#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/r02w_bench_8gpu.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02w_bench_8gpu.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d["value"], d["e2e"]["value"], d["n_gpus"], d["scaling"], d["ms_per_step"], d["parity"]["ok"] if d.get("parity") else None, d["config"].get("shard_imbalance"))
else: print(open('gpurun_out/r02w_bench_8gpu.log').read()[-2500:])
PY
