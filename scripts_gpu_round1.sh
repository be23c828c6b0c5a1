#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/bench.log | head -2; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench.log; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --rows 1250000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.log 2> gpurun_out/bench_s.err; echo "bench 1.25M rc=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_s.log; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_s.log
