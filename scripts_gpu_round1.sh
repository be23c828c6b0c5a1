#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor_u8.py -m gpu -x -q > gpurun_out/pytest_u8.log 2>&1; echo "pytest u8 rc=$?"
tail -30 gpurun_out/pytest_u8.log | cut -c1-300
