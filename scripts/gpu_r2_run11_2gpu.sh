#!/bin/bash
# 2 GPUs: NCCL path of the shard group (one process / two devices, and one process per GPU under torchrun)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_2gpu.log | cut -c1-600
for n in 1 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err; echo "bench n=$n rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n$n.log") if l.startswith("{")][-1])
    print("n=$n value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],3),"e2e ms",round(d["e2e"]["ms_per_step"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"frac",round(d["roofline"]["frac"],3),"launches",d["gpu_launches"], d["parity_checked"])
except Exception as e: print("no line", e)
PY
tail -4 gpurun_out/bench_n$n.err | cut -c1-300
done
