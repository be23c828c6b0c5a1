#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "pytest hnsw+fullsize rc=$?"; tail -12 gpurun_out/pytest_a.log | cut -c1-300
timeout 900 python bench.py --workload hnsw --steps 5 --warmup 3 > gpurun_out/bench_hnsw.log 2> gpurun_out/bench_hnsw.err; echo "hnsw rc=$?"; cut -c1-700 gpurun_out/bench_hnsw.log; grep -o '"roofline.*' gpurun_out/bench_hnsw.log | cut -c1-900
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "brute_raw_small_shapes and 33 or score_ids or rerank or quaternary_1024 or distance_pairs and 33" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -6 gpurun_out/sanitizer.log | cut -c1-300
