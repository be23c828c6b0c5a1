#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hnsw.py -m gpu -q > gpurun_out/pytest_hnsw.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_hnsw.log | cut -c1-300
rows=1000000
timeout 900 python bench.py --workload c3 --rows $rows --steps 5 --warmup 3 --hnsw-variants > gpurun_out/c3_${rows}_variants.json 2> gpurun_out/c3_${rows}_variants.err; echo "c3 rows=$rows rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/c3_${rows}_variants.json"))
print("QPS",round(d["value"]),"e2e",round(d["e2e"]["value"]),"recall",d["recall_at_10"],"frac",round(d["roofline"]["frac"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"build_s",round(d["build_seconds"],1), d["clocks"])
for v in d["hnsw_variants"][:4]:
    print(f'{v["variant"]:28s} B={v["batch"]:5d} {v["kernel_ms"]:.3f} ms {v["kernel_qps"]/1000:.1f}k eq={v["ids_equal_default"]}', v["cycles_per_pop"])
PY
tail -n 2 gpurun_out/c3_${rows}_variants.err
timeout 900 python bench.py --workload c3 --rows 10000000 --storage bf16 --steps 10 --warmup 3 > gpurun_out/c3_10M_bf16.json 2> gpurun_out/c3_10M_bf16.err; echo "c3 bf16 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/c3_10M_bf16.json"))
print("bf16 QPS",round(d["value"]),"e2e",round(d["e2e"]["value"]),"recall",d["recall_at_10"],"frac",round(d["roofline"]["frac"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"build_s",round(d["build_seconds"],1), d["clocks"])
PY
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/c4_50M.json 2> gpurun_out/c4_50M.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/c4_50M.json"))
print("c4 QPS",round(d["value"]),"ms",round(d["ms_per_step"],2),"frac",round(d["roofline"]["frac"],3), d["roofline"].get("peak"), d["clocks"])
PY
tail -n 2 gpurun_out/c4_50M.err
