#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --workload c3 --rows 10000000 --steps 5 --warmup 3 > gpurun_out/bench_c3_10m.log 2> gpurun_out/bench_c3_10m.err; echo "c3 10M rc=$?"; cut -c1-2600 gpurun_out/bench_c3_10m.log; tail -5 gpurun_out/bench_c3_10m.err
