#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02v_pytest_gpu.txt 2>&1; tail -12 gpurun_out/r02v_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.txt 2>&1; tail -2 gpurun_out/r02v_smoke.txt
python bench.py > gpurun_out/r02v_bench_full.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02v_bench_full.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d["value"], d["e2e"]["value"], d["components"]["analysis"]["audio_s_per_s"], d["components"]["synthesis"]["audio_s_per_s"], d["parity"]["ok"], d["roofline"]["kernel"], d["roofline"]["frac"])
    print({k:(v["share_of_step"], v["avg_launch_ms"], v["hbm_frac"]) for k,v in d["kernels"].items()})
    print({k:(v["value"], v["e2e"]["value"], v["roofline"]["frac"]) for k,v in d["workloads"].items()})
else:
    print(open('gpurun_out/r02v_bench_full.log').read()[-2500:])
PY
