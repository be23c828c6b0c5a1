#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_tensor_path.py tests/test_gpu_metadata.py tests/test_gpu_hnsw_build.py -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log | cut -c1-600
timeout 600 python bench.py --workload c3 --rows 1000000 --steps 5 --warmup 3 --hnsw-prof > gpurun_out/c3_1M_warp_v2.json 2> gpurun_out/c3_1M_warp_v2.err; echo "c3 rc=$?"; cat gpurun_out/c3_1M_warp_v2.json; tail -3 gpurun_out/c3_1M_warp_v2.err
