#!/bin/bash
mkdir -p gpurun_out
for v in "" rb3 rb6; do
  if [ -n "$v" ]; then export B2W_LIB=variants/libb200world_$v.so; fi
  python bench.py --utts 1024 --steps 2 --warmup 1 --no-workloads --no-cpu-baseline > gpurun_out/r02l_bench_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/r02l_bench_$v.log').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("$v", d["components"]["synthesis"]["audio_s_per_s"], d["parity"]["ok"], d["kernels"]["render"]["avg_launch_ms"])
else: print(open('gpurun_out/r02l_bench_$v.log').read()[-1000:])
PY
done
