#!/bin/bash
# one gpurun call: tensor-path tests first (short timeout), full parity, smoke, bench, ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor_path.py -m gpu -x -q > gpurun_out/pytest_tensor.log 2>&1; echo "pytest tensor rc=$?"
tail -25 gpurun_out/pytest_tensor.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --exact-only --no-cpu-baseline > gpurun_out/bench_exact.log 2> gpurun_out/bench_exact.err; echo "bench exact rc=$?"; cat gpurun_out/bench_exact.log
timeout 600 python bench.py --batch 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b8.log 2> gpurun_out/bench_b8.err; echo "bench8 rc=$?"; cat gpurun_out/bench_b8.log; tail -3 gpurun_out/bench_b8.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
grep -E "tensor_scan|rerank|scan_f32" gpurun_out/launches.csv | tail -8
