"""Steady-state A/B of K2 variants on ONE box (config-3 shape): blocks of back-to-back launches, the
variants interleaved (A B A B ...), median of the second half of every block, SM clock and board power
sampled through NVML while the block runs.  Variants are "name:ENV=VALUE,ENV=VALUE".

    python profiles/k2_ab.py "classic:QST_SCORE_QS=0" "qs:QST_SCORE_QS=1" [--blocks 3] [--launches 40]
"""
import ctypes as C
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402
from qst_b200 import _lib, scoring  # noqa: E402

args, opts, it = [], {}, iter(sys.argv[1:])
for a in it:
    if a.startswith("--"):
        opts[a] = next(it)
    else:
        args.append(a)
BLOCKS = int(opts.get("--blocks", 3))
LAUNCHES = int(opts.get("--launches", 40))
Q, N, D, K = int(opts.get("--q", 10_000)), int(opts.get("--n", 1_000_000)), int(opts.get("--d", 768)), 100
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(14)
corpus = torch.cat([torch.randn(N // 8, D, generator=g, device=dev) for _ in range(8)])
queries = torch.randn(Q, D, generator=g, device=dev)
index = qst_b200.CorpusIndex(corpus, "cos_sim")
del corpus
pq = scoring.prepare_rows(queries, True)
lib = _lib.load()
st = _lib.stream_ptr(dev)

try:
    import pynvml
    pynvml.nvmlInit()
    H = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:
    pynvml = None


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.mhz, self.watts, self.halt = [], [], threading.Event()

    def run(self):
        while pynvml is not None and not self.halt.is_set():
            try:
                self.mhz.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
                self.watts.append(pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0)
            except Exception:
                pass
            self.halt.wait(0.01)


variants = []
for a in args:
    name, _, envs = a.partition(":")
    variants.append((name, dict(e.split("=") for e in envs.split(",") if e)))
KEYS = sorted({k for _, e in variants for k in e})


def set_env(env):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)


results = {name: [] for name, _ in variants}
for b in range(BLOCKS):
    for name, env in variants:
        set_env(env)
        plan = scoring.make_plan(Q, N, D, K, 0, "cos_sim")
        ws = scoring._workspace(plan.ws_bytes, dev, "select")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(LAUNCHES + 1)]
        smp = Sampler()
        smp.start()
        ev[0].record()
        for i in range(LAUNCHES):
            _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), index.rows.bf16.data_ptr(), ws.data_ptr(), st))
            ev[i + 1].record()
        torch.cuda.synchronize()
        smp.halt.set()
        smp.join()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(LAUNCHES // 2, LAUNCHES))
        med = ts[len(ts) // 2]
        mhz = sorted(smp.mhz[len(smp.mhz) // 2:])
        w = sorted(smp.watts[len(smp.watts) // 2:])
        results[name].append(med)
        print(f"block {b} {name:>14s} qs={plan.qs} ctas={plan.ctas}: median {med:.3f} ms -> {2 * Q * N * D / med / 1e9:.0f} TFLOP/s  "
              f"first {ev[0].elapsed_time(ev[1]):.3f} ms  sm {mhz[len(mhz) // 2] if mhz else '?'} MHz  "
              f"{w[len(w) // 2] if w else 0:.0f} W", flush=True)
print("summary (median over blocks of the steady-state medians):")
for name, v in results.items():
    v = sorted(v)
    print(f"  {name:>14s}: {v[len(v) // 2]:.3f} ms -> {2 * Q * N * D / v[len(v) // 2] / 1e9:.0f} TFLOP/s")
