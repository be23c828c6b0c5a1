#!/bin/bash
# 8 GPUs of one box: the sharded tests on real NCCL, the headline at N = 8 and 4 (strong scaling), configs[4] (c5) at 8 x 2M rows
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/pytest_8gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_8gpu.log | cut -c1-300
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err; echo "bench n=$n rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/bench_n$n.log"):
    if ln.startswith("{"):
        d=json.loads(ln)
        print("n=$n value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"frac",round(d["roofline"]["frac"],3),"launches",d["gpu_launches"], d.get("parity_checked"), d["clocks"])
PY
tail -n 2 gpurun_out/bench_n$n.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --workload c5 --gpus 8 --rows 16000000 --steps 10 --warmup 3 > gpurun_out/c5_8gpu.log 2> gpurun_out/c5_8gpu.err; echo "c5 rc=$?"
python - <<'PY'
import json
for ln in open("gpurun_out/c5_8gpu.log"):
    if ln.startswith("{"):
        d=json.loads(ln)
        print("c5 value",round(d["value"]),"ms/step",round(d["ms_per_step"],3),"recall",d["recall_at_10"],"frac",round(d["roofline"]["frac"],3),"build",d["build_seconds_max_over_ranks"], d["config"]["workload"][:90])
PY
tail -n 3 gpurun_out/c5_8gpu.err
