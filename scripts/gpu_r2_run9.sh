#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest.log | cut -c1-800
cat > gpurun_out/c1.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, cosdata_b200 as cdb
ix = cdb.DenseIndex(dim=128, capacity=100000); ix.append_synthetic(7, 100000)
q = cdb.synth_matrix(8, 1, 128)
for _ in range(6): ix.batch_search(q, 10)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c1.csv python gpurun_out/c1.py > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_c1.csv')) if len(r)>5]
hdr=[r for r in rows if 'Kernel Name' in r][0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[rows.index(hdr)+1:][-12:]: print(r[ki].split('(')[0][:60], r[vi])
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_quick.log"))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"frac",round(d["roofline"]["frac"],3),"launches",d["gpu_launches"])
for s in d.get("secondary",[]):
    print(" ", s.get("name","?")[:70], "|", round(s.get("value",0),1), s.get("unit"), "ms", round(s.get("ms_per_step",0),4), "cold", s.get("cold_l2_ms_per_step"), "frac", round(s.get("roofline",{}).get("frac",0),3), s.get("error",""))
PY
tail -3 gpurun_out/bench_quick.err
