#!/bin/bash
mkdir -p gpurun_out
nproc
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest.log | cut -c1-800
( time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err ) 2>&1 | grep real; echo "bench rc=$?"; cat gpurun_out/bench.log | cut -c1-6000; tail -5 gpurun_out/bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; cat gpurun_out/bench_ref.log | cut -c1-1500
