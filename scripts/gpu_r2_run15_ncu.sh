#!/bin/bash
# ncu evidence for the round-2 kernels (one GPU; numbers printed under ncu are never bench values)
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of the headline command
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2_launches_c2_b1024.csv \
  python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
# 2. tensor_scan_kernel, cta_group::2 form, 10M x 768, B = 1024: two consecutive launches (seeding pass + main pass)
timeout 1200 $NCU --set full --import-source on -k regex:tensor_scan_kernel --launch-skip 6 --launch-count 2 -f -o gpurun_out/r2_tensor_scan_pair_b1024 \
  python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo "tensor_scan rc=$?"
# 3. hnsw_search_warp_kernel, 1M x 768 f16, ef 128, B = 1024
timeout 1200 $NCU --set full --import-source on -k regex:hnsw_search_warp_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_hnsw_warp_1M_768 \
  python bench.py --workload c3 --rows 1000000 --steps 1 --warmup 3 > gpurun_out/ncu3.log 2>&1; echo "hnsw rc=$?"
# 4. tensor_scan_kernel kind::i8, quaternary 5M x 1024, B = 2048 (one chunk of C4)
cat > gpurun_out/c4_probe.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch, cosdata_b200 as cdb
r, D, B, k = 5_000_000, 1024, 2048, 10
ix = cdb.DenseIndex(dim=D, storage_type=cdb.StorageType.SubByte2, metric=cdb.DistanceMetricKind.DotProduct, capacity=r)
ix.append_synthetic(0xC05DA7A + 4, r)
q = torch.from_numpy(cdb.synth_matrix(0xC05DA7A + 104, B, D)).cuda()
i = torch.empty((B, k), dtype=torch.int32, device="cuda"); s = torch.empty((B, k), dtype=torch.float32, device="cuda")
st = torch.cuda.Stream()
for _ in range(4):
    ix.batch_search_device(q.data_ptr(), B, k, i.data_ptr(), s.data_ptr(), None, None, st.cuda_stream, mode=cdb.SearchMode.BRUTE_CODES)
torch.cuda.synchronize()
print(ix.stats())
PY
timeout 1200 $NCU --set full --import-source on -k regex:tensor_scan_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r2_tensor_scan_i8_5M \
  python gpurun_out/c4_probe.py > gpurun_out/ncu4.log 2>&1; echo "i8 rc=$?"
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/ncu2.log gpurun_out/ncu3.log gpurun_out/ncu4.log | cut -c1-300
