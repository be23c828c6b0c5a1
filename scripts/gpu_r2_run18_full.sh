#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_full.log | cut -c1-400
rows=1000000
timeout 900 python bench.py --workload c3 --rows $rows --steps 5 --warmup 3 --hnsw-variants > gpurun_out/c3_${rows}_variants.json 2> gpurun_out/c3_${rows}_variants.err; echo "c3 rows=$rows rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/c3_${rows}_variants.json"))
print("QPS",round(d["value"]),"e2e",round(d["e2e"]["value"]),"recall",d["recall_at_10"],"frac",round(d["roofline"]["frac"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"build_s",round(d["build_seconds"],1), d["clocks"])
for v in d["hnsw_variants"][:4]:
    print(f'{v["variant"]:28s} B={v["batch"]:5d} {v["kernel_ms"]:.3f} ms {v["kernel_qps"]/1000:.1f}k eq={v["ids_equal_default"]}', v["cycles_per_pop"])
PY
tail -n 2 gpurun_out/c3_${rows}_variants.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_full.log"))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"frac",round(d["roofline"]["frac"],3),"launches",d["gpu_launches"], d.get("cpu_baseline"))
for s in d.get("secondary",[]):
    print(" ", s.get("name","?")[:70], "|", round(s.get("value",0),1), s.get("unit"), "ms", round(s.get("ms_per_step",0),4), "cold", s.get("cold_l2_ms_per_step"), "frac", round(s.get("roofline",{}).get("frac",0),3), "recall", s.get("recall_at_10"), s.get("error",""))
PY
tail -n 3 gpurun_out/bench_full.err
