#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor_path.py -m gpu -x -q > gpurun_out/pytest_tensor.log 2>&1; echo "pytest tensor rc=$?"
tail -4 gpurun_out/pytest_tensor.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tensor_scan -s 3 -c 1 -o gpurun_out/prof_tensor python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tensor.log 2>&1; echo "ncu tensor rc=$?"
