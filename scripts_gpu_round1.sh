#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor_path.py -m gpu -x -q > gpurun_out/pytest_tensor.log 2>&1; echo "pytest tensor rc=$?"
tail -12 gpurun_out/pytest_tensor.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --batch 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b8.log 2> gpurun_out/bench_b8.err; echo "bench8 rc=$?"; cat gpurun_out/bench_b8.log; tail -3 gpurun_out/bench_b8.err
timeout 600 python bench.py --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b64.log 2> gpurun_out/bench_b64.err; echo "bench64 rc=$?"; cat gpurun_out/bench_b64.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tensor_scan -s 3 -c 1 -o gpurun_out/prof_tensor python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tensor.log 2>&1; echo "ncu tensor rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_f32 -s 1 -c 1 -o gpurun_out/prof_scan_b8 python bench.py --batch 8 --exact-only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_scan.log 2>&1; echo "ncu scan rc=$?"
ls -la gpurun_out/*.ncu-rep
