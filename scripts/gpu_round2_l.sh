#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synthesis.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/r02s_pytest.txt 2>&1; tail -3 gpurun_out/r02s_pytest.txt
python bench.py --utts 2048 --steps 2 --warmup 1 --no-workloads --no-cpu-baseline > gpurun_out/r02s_bench.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/r02s_bench.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d["value"], d["e2e"]["value"], d["components"]["synthesis"]["audio_s_per_s"], d["parity"]["ok"], d["parity"]["resynthesis_snr_db_min"], {k:v["avg_launch_ms"] for k,v in d["kernels"].items()})
else: print(open('gpurun_out/r02s_bench.log').read()[-1500:])
PY
