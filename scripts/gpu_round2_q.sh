#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_distributed.py -m gpu -q > gpurun_out/r02u_pytest.txt 2>&1; tail -4 gpurun_out/r02u_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-workloads > gpurun_out/r02u_bench_2gpu.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02u_bench_2gpu.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d["value"], d["e2e"]["value"], d["n_gpus"], d["scaling"], d["parity"]["ok"] if d.get("parity") else None)
else: print(open('gpurun_out/r02u_bench_2gpu.log').read()[-2000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02u_bench_ref_2gpu.log 2>&1; tail -c 600 gpurun_out/r02u_bench_ref_2gpu.log
