#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor_path.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_t.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_t.log | cut -c1-300
for r in 1250000 10000000; do
timeout 300 python bench.py --rows $r --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r$r.log 2>gpurun_out/bench_r$r.err; grep -o '"value": [0-9.]*' gpurun_out/bench_r$r.log | head -1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_r$r.log | head -1; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_r$r.log; grep -o '"candidates_per_query": {[^}]*}' gpurun_out/bench_r$r.log
done
