#!/bin/bash
# Round-end style validation on a B200 box (run under gpurun): parity tests, smoke, headline bench, reference arm.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_validate.sh'
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.log
