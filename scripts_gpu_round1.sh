#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hnsw_build.py tests/test_gpu_hnsw.py -m gpu -x -q > gpurun_out/pytest_build.log 2>&1; echo "pytest build rc=$?"; tail -25 gpurun_out/pytest_build.log | cut -c1-400
