#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hnsw_build.py tests/test_gpu_metadata.py tests/test_gpu_hnsw.py -m gpu -q -x > gpurun_out/pytest_replicas.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_replicas.log | cut -c1-400
