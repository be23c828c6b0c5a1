#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor_path.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_t.log | cut -c1-300
run() { timeout 300 python bench.py --rows $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.log 2>gpurun_out/b.err; echo "$1: $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/b.log) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/b.log | head -1) $(grep -o '"candidates_per_query": {[^}]*}' gpurun_out/b.log)"; }
CDB_TS_REFRESH=7 run "refresh7" 10000000; run "default" 10000000
CDB_TS_REFRESH=15 run "refresh15" 10000000
CDB_TS_REFRESH=31 run "refresh31" 10000000
CDB_TS_REFRESH=15 run "refresh15 1.25M" 1250000
