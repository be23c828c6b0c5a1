#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hnsw.py -m gpu -q > gpurun_out/pytest_hnsw.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_hnsw.log | cut -c1-300
rows=1000000
timeout 900 python bench.py --workload c3 --rows $rows --steps 5 --warmup 3 --hnsw-variants > gpurun_out/c3_${rows}_variants.json 2> gpurun_out/c3_${rows}_variants.err; echo "c3 rows=$rows rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/c3_${rows}_variants.json"))
print("QPS",round(d["value"]),"e2e",round(d["e2e"]["value"]),"recall",d["recall_at_10"],"frac",round(d["roofline"]["frac"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"build_s",round(d["build_seconds"],1), d["clocks"])
for v in d["hnsw_variants"][:4]:
    print(f'{v["variant"]:28s} B={v["batch"]:5d} {v["kernel_ms"]:.3f} ms {v["kernel_qps"]/1000:.1f}k eq={v["ids_equal_default"]}', v["cycles_per_pop"])
PY
tail -2 gpurun_out/c3_${rows}_variants.err
bash scripts/gpu_r2_run15_ncu.sh
# the speculative variant under ncu as well
timeout 1200 ncu --clock-control none --set full --import-source on -k regex:hnsw_search_warp_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_hnsw_warp_spec_1M_768 \
  python bench.py --workload c3 --rows 1000000 --steps 1 --warmup 3 --hnsw-flags 38 > gpurun_out/ncu5.log 2>&1; echo "hnsw spec rc=$?"
ls -la gpurun_out/*.ncu-rep
