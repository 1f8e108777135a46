#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_analysis.py tests/test_gpu_synthesis.py tests/test_gpu_pipeline.py -m gpu -q > gpurun_out/r02i_pytest.txt 2>&1; tail -12 gpurun_out/r02i_pytest.txt
python scripts/gpu_kbench.py --utts 512 --kernels cheaptrick,mcep > gpurun_out/r02i_kbench.txt 2>&1; cat gpurun_out/r02i_kbench.txt
python bench.py --utts 600 --steps 2 --warmup 1 --no-workloads --no-cpu-baseline > gpurun_out/r02i_bench_small.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02i_bench_small.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d["value"], d["components"]["analysis"]["audio_s_per_s"], d["components"]["synthesis"]["audio_s_per_s"], d["parity"])
    print({k:(v["share_of_step"], v["avg_launch_ms"]) for k,v in d["kernels"].items()})
else:
    print(open('gpurun_out/r02i_bench_small.log').read()[-1500:])
PY
