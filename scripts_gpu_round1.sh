#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 2 -c 1 -o gpurun_out/prof_hnsw python bench.py --workload c3 --rows 1000000 --steps 1 --warmup 1 > gpurun_out/ncu_hnsw.log 2>&1; echo "ncu hnsw rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tensor_scan_u8 -s 5 -c 1 -o gpurun_out/prof_u8 python bench.py --workload c4 --rows 5000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_u8.log 2>&1; echo "ncu u8 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep
