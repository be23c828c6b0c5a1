#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor_u8.py tests/test_gpu_tensor_path.py tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest.log | cut -c1-800
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_quick.log"))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"frac",round(d["roofline"]["frac"],3),"launches",d["gpu_launches"])
for s in d.get("secondary",[]):
    print(" ", s.get("name","?")[:70], "|", round(s.get("value",0),1), s.get("unit"), "ms", round(s.get("ms_per_step",0),4), "cold", s.get("cold_l2_ms_per_step"), "frac", round(s.get("roofline",{}).get("frac",0),3), s.get("error",""))
PY
tail -3 gpurun_out/bench_quick.err
