#!/bin/bash
# Full verification: all GPU tests, smoke, the default bench line (with every workload).
tag=${1:-r02z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.txt 2>&1; tail -6 gpurun_out/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1; tail -2 gpurun_out/${tag}_smoke.txt
python bench.py > gpurun_out/${tag}_bench_full.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/${tag}_bench_full.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d["value"], d["e2e"]["value"], d["components"]["analysis"]["audio_s_per_s"], d["components"]["synthesis"]["audio_s_per_s"], d["parity"]["ok"])
    w=d["workloads"]; print(w["synth256"]["value"], w["vtln109"]["value"]); print(json.dumps(w["postprocess256"], indent=1)); print(json.dumps(w["gen_data_files"], indent=1))
else:
    print(open('gpurun_out/${tag}_bench_full.log').read()[-2500:])
PY
