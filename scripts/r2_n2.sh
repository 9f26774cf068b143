#!/bin/bash
# 2 GPUs: real-NCCL test of the sharded path (C-ABI communicator, copy-engine gathers, prefetch, fused exchanges) + A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_n2_tests.txt 2>&1
tail -15 gpurun_out/r2_n2_tests.txt | cut -c1-300
export QST_BENCH_SKIP_SECONDARY=1
for rep in 1 2; do
for nf in 1 ""; do
  QST_NO_FUSED_EXCHANGE=$nf timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ab_n2_nf${nf}_$rep.json 2> gpurun_out/r2_ab_n2.err
  echo "bench rc=$? nofused=$nf"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_ab_n2_nf${nf}_$rep.json"))
print("AB nofused='${nf}'", round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), d["parity_sample"]["mismatch"], {k:round(v,3) for k,v in d.get("stage_ms_rank0",{}).items()})
PY
done
done
tail -c 1500 gpurun_out/r2_ab_n2.err
