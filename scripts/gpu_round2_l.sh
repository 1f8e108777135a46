#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synthesis.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/r02s_pytest.txt 2>&1; tail -3 gpurun_out/r02s_pytest.txt
timeout 600 python scripts/gpu_synth_e2e_modes.py > gpurun_out/r02s_synth_modes.txt 2>&1; cat gpurun_out/r02s_synth_modes.txt
python bench.py --no-cpu-baseline > gpurun_out/r02s_bench.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/r02s_bench.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d["value"], d["e2e"]["value"], d["e2e_components"], d["components"], d["parity"]["ok"], {k:v["avg_launch_ms"] for k,v in d["kernels"].items()}, {k:(v["value"], v["e2e"]["value"]) for k,v in d["workloads"].items()})
else: print(open('gpurun_out/r02s_bench.log').read()[-1500:])
PY
