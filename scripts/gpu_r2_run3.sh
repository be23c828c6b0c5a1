#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.log | cut -c1-600
for c in 1 0; do
CDB_HNSW_CONV=$c timeout 600 python bench.py --workload c3 --rows 1000000 --steps 5 --warmup 3 --hnsw-prof > gpurun_out/c3_1M_warp_v3_conv$c.json 2> gpurun_out/c3_1M_warp_v3_conv$c.err; echo "c3 conv=$c rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/c3_1M_warp_v3_conv$c.json"))
print("QPS",round(d["value"]),"e2e",round(d["e2e"]["value"]),"recall",d["recall_at_10"],"frac",round(d["roofline"]["frac"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3))
print({k:round(v) for k,v in d["hnsw_phase_profile"]["cycles_per_pop"].items()})
PY
tail -2 gpurun_out/c3_1M_warp_v3_conv$c.err
done
