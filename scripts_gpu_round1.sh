#!/bin/bash
# one gpurun call: parity tests, smoke, bench (B=1024 and B=8), ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --batch 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b8.log 2> gpurun_out/bench_b8.err; echo "bench8 rc=$?"; cat gpurun_out/bench_b8.log; tail -3 gpurun_out/bench_b8.err
timeout 600 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b1.log 2> gpurun_out/bench_b1.err; echo "bench1 rc=$?"; cat gpurun_out/bench_b1.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/launches.csv
