#!/bin/bash
# cta_group::2 tensor scan: parity first (small shapes, short timeouts), then A/B timing against the single-CTA form
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor_path.py tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/pytest_pair.log 2>&1; echo "pytest(pair) rc=$?"; tail -6 gpurun_out/pytest_pair.log | cut -c1-600
CDB_TS_PAIR=0 timeout 300 python -m pytest tests/test_gpu_tensor_path.py -m gpu -q -x > gpurun_out/pytest_single.log 2>&1; echo "pytest(single) rc=$?"; tail -3 gpurun_out/pytest_single.log | cut -c1-300
for pair in 1 0; do
CDB_TS_PAIR=$pair timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_pair$pair.log 2> gpurun_out/bench_pair$pair.err; echo "bench pair=$pair rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_pair$pair.log"))
    print("pair=$pair value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"frac",round(d["roofline"]["frac"],3),"cand",d["roofline"]["candidates_per_query"],d["clocks"])
except Exception as e: print("no line", e)
PY
tail -3 gpurun_out/bench_pair$pair.err
done
# per-kernel launch list of the C1 shape (100k x 128, batch 1): where do 0.45 ms go?
cat > /tmp/c1.py <<'PY'
import numpy as np, cosdata_b200 as cdb
ix = cdb.DenseIndex(dim=128, capacity=100000); ix.append_synthetic(7, 100000)
q = cdb.synth_matrix(8, 1, 128)
for _ in range(6): ix.batch_search(q, 10)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c1.csv python /tmp/c1.py > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_c1.csv')) if len(r)>5]
hdr=[r for r in rows if 'Kernel Name' in r][0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[rows.index(hdr)+1:][-14:]: print(r[ki].split('(')[0][:60], r[vi])
PY
