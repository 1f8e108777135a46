#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_metrics.py tests/test_mlpg.py tests/test_gpu_labels.py tests/test_gpu_pipeline.py tests/test_gpu_batching.py -m gpu -x -q > gpurun_out/r03a_pytest_post.txt 2>&1; tail -4 gpurun_out/r03a_pytest_post.txt
python bench.py --utts 2048 --steps 5 --warmup 3 --no-cpu-baseline --parity-utts 0 --io-utts 0 > gpurun_out/r03a_bench_small.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r03a_bench_small.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(json.dumps(d["workloads"]["postprocess256"]["operators"], indent=1))
else:
    print(open('gpurun_out/r03a_bench_small.log').read()[-2500:])
PY
