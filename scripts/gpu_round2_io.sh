#!/bin/bash
# N4 file IO: the gen_data / reader tests on the GPU, then the bench with the gen_data_files workload.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dio.py tests/test_corpus_io.py -m "gpu or not gpu" -x -q > gpurun_out/r02x_pytest_io.log 2>&1
tail -3 gpurun_out/r02x_pytest_io.log
python bench.py > gpurun_out/r02x_bench_full.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02x_bench_full.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['e2e']['value'], json.dumps(d['workloads']['gen_data_files'], indent=1))
else:
    print(open('gpurun_out/r02x_bench_full.log').read()[-3000:])
PY
