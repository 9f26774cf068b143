#!/bin/bash
# one 8-GPU box: N = 8 full line (prefetch + fused exchanges), then A/B variants (headline block only), then N = 1
set -x
mkdir -p gpurun_out
run8() { # name, env...
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 200)) bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"
}
run8 r2c_scale_n8 X=1
run8 r2c_n8_plain QST_BENCH_SKIP_SECONDARY=1 QST_NO_FUSED_EXCHANGE=1 QST_BENCH_NO_PREFETCH=1
run8 r2c_n8_prefetch_only QST_BENCH_SKIP_SECONDARY=1 QST_NO_FUSED_EXCHANGE=1
run8 r2c_n8_fused_only QST_BENCH_SKIP_SECONDARY=1 QST_BENCH_NO_PREFETCH=1
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_scale_n1.json 2> gpurun_out/r2c_scale_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
for f in ("r2c_scale_n1", "r2c_scale_n8", "r2c_n8_plain", "r2c_n8_prefetch_only", "r2c_n8_fused_only"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["ms_per_step"], 3), round(d["value"]), round(d["e2e"]["value"]), d["parity_sample"]["mismatch"],
              {k: round(v, 3) for k, v in (d.get("stage_ms_rank0") or {}).items()},
              {k: (round(v.get("value", 0)) if isinstance(v, dict) else v) for k, v in d.items() if k in ("config4", "config5", "replicated_master")})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -c 600 gpurun_out/r2c_scale_n8.err
