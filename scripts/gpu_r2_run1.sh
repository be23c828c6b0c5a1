#!/bin/bash
# round-2 run 1: parity tests on the new paths + HNSW kernel A/B (CTA-per-query vs warp-per-query) at 1M x 768 f16
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log | cut -c1-400
for v in 1 0; do
  CDB_HNSW_CTA=$v timeout 600 python bench.py --workload c3 --rows 1000000 --steps 5 --warmup 3 --hnsw-prof --dump-ids gpurun_out/hnsw_ids_cta$v.npz > gpurun_out/c3_1M_cta$v.json 2> gpurun_out/c3_1M_cta$v.err; echo "c3 cta=$v rc=$?"; cat gpurun_out/c3_1M_cta$v.json; tail -3 gpurun_out/c3_1M_cta$v.err
done
python - <<'PY'
import numpy as np
a=np.load('gpurun_out/hnsw_ids_cta1.npz'); b=np.load('gpurun_out/hnsw_ids_cta0.npz')
print("A/B ids equal:", np.array_equal(a['ids'],b['ids']), "scores equal:", np.array_equal(a['scores'].view(np.uint32), b['scores'].view(np.uint32)))
PY
