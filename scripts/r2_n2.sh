#!/bin/bash
# 2 GPUs: real-NCCL test of the sharded path (incl. the C-ABI communicator, copy-engine gathers, prefetch) + A/B of prefetch
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_n2_tests.txt 2>&1
tail -15 gpurun_out/r2_n2_tests.txt
export QST_BENCH_SKIP_SECONDARY=1
for rep in 1 2; do
for np in 1 ""; do
  QST_BENCH_NO_PREFETCH=$np timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ab_n2_np${np}_$rep.json 2> gpurun_out/r2_ab_n2.err
  echo "bench rc=$? noprefetch=$np"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_ab_n2_np${np}_$rep.json"))
print("AB noprefetch='${np}'", d["ms_per_step"], d["value"], d["e2e"]["value"], d["parity_sample"], d.get("stage_ms_rank0"))
PY
done
done
tail -c 1500 gpurun_out/r2_ab_n2.err
