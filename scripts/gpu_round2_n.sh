#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vtln.py -m gpu -q -x > gpurun_out/r02n_pytest.txt 2>&1; tail -4 gpurun_out/r02n_pytest.txt
timeout 300 python scripts/gpu_vtln_bench.py bwd > gpurun_out/r02n_vtln.txt 2>&1; cat gpurun_out/r02n_vtln.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02n_vtln_launches.csv python scripts/gpu_vtln_bench.py bwd > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02n_vtln_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:60]].append(float(r[vi].replace(',','')))
for k,v in agg.items(): print("%-62s n=%3d  median %.1f us  min %.1f" % (k, len(v), sorted(v)[len(v)//2]/1e3, min(v)/1e3))
PY
