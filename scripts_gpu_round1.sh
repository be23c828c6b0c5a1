#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus=$N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --workload c5 --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"; cut -c1-2200 gpurun_out/bench_c5.log; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_c5.err | tail -5 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_c2_${N}gpu.log 2> gpurun_out/bench_c2_${N}gpu.err; echo "c2 x$N rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/bench_c2_${N}gpu.log | head -1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c2_${N}gpu.log | head -1
