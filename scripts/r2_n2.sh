#!/bin/bash
# 2 GPUs: real-NCCL test of the sharded path (C-ABI communicator, copy-engine gathers, prefetch, fused exchanges) + the N=2 line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_n2_tests.txt 2>&1
tail -5 gpurun_out/r2_n2_tests.txt | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_scale_n2.json 2> gpurun_out/r2c_scale_n2.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c_scale_n2.json"))
print("N2", round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), d["parity_sample"]["mismatch"], {k:round(v,3) for k,v in d.get("stage_ms_rank0",{}).items()},
      {k: (round(v.get("value", 0)) if isinstance(v, dict) else v) for k, v in d.items() if k in ("config4", "config5", "replicated_master")})
PY
tail -c 500 gpurun_out/r2c_scale_n2.err
