"""bench.py's JSON line without a GPU: the expression that builds the line in ``run_ours`` is lifted out with
``ast`` and evaluated with stand-in measurements, so a typo in it (the one part of the bench no CPU box
executes) shows up here; the reference arm's line is built by the real function with the timing stubbed."""
import ast
import importlib.util
import json
import os
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "parity_sample"]


def _ours_line(bench, **over):
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_ours")
    assign = next(n for n in ast.walk(fn) if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "line")
    ns = dict(vars(bench))
    ns.update(Q=10000, ms_step=11.4, world=1, steps=10, warmup=3, uncertified=0, ms_e2e=12.0, ms_e2e_serial=12.9,
              h2d=10000 * 768 * 4, d2h=10000 * 100 * 12, e2e_mode="double-buffered", pipe_same=True, launches=90,
              achieved=1500.0, peak=1401.5, traffic=2.7e9, kernel_name="score_select_qs_kernel", k2=10.1,
              peaks={"_source": "measured", "bf16_tflops": 1655.8}, parity={"mismatch": 0}, clocks={"sm_mhz": 1400},
              plan=types.SimpleNamespace(kprime=224, m_tiles=40, n_tiles=3907, stripes=46, units=1840, grid=74, qs=1),
              index=types.SimpleNamespace(n=1_000_000))
    ns.update(over)
    return eval(compile(ast.Expression(body=assign.value), "bench.py", "eval"), ns)


def test_our_arms_line_has_the_contract_keys_and_consistent_arithmetic():
    bench = _bench()
    line = _ours_line(bench)
    json.dumps(line)
    for key in REQUIRED:
        assert key in line, key
    assert line["value"] == 10000 / 11.4e-3 and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["vs_baseline"] is None and line["dtype"] == "bf16" and line["data"] == "synthetic"
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert line["roofline"]["frac"] == 1500.0 / 1401.5 and line["roofline"]["bound"] == "tensor"
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["config"]["workload"] == bench.WORKLOAD and "model" not in line["config"]
    # e2e headline = the faster of the two timed host-buffer entries, both kept
    assert line["e2e"]["value"] == line["e2e"]["pipelined_value"] == 10000 / 12.0e-3
    slow = _ours_line(bench, ms_e2e=15.2, ms_e2e_serial=12.7)
    assert slow["e2e"]["value"] == slow["e2e"]["serial_call_value"] == 10000 / 12.7e-3
    assert slow["e2e"]["mode"].startswith("one synchronous") and slow["e2e"]["pipelined_ms_per_step"] == 15.2


def test_reference_arms_line_names_the_same_metric_and_workload(monkeypatch, capsys):
    bench = _bench()
    fake = {"value": 385.0, "unit": "queries/s", "cores": 16, "kind": "port", "sample": "Q=1000 x full corpus",
            "steps": 10, "warmup": 3, "ms_per_sample_step": 2597.0}
    monkeypatch.setattr(bench, "cpu_reference_run", lambda steps, warmup, budget_s=150.0: fake)
    monkeypatch.setattr(bench, "emit", lambda line: print(json.dumps(line)))
    monkeypatch.delenv("RANK", raising=False)
    bench.run_reference(types.SimpleNamespace(gpus=1, steps=10, warmup=3))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    ours = _ours_line(bench)
    assert line["impl"] == "reference" and line["gpu_launches"] == 0
    for key in ("metric", "unit", "higher_is_better", "scaling"):
        assert line[key] == ours[key], key
    assert line["config"]["workload"] == ours["config"]["workload"]
    assert line["e2e"] == {"value": 385.0, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"] == 385.0
    assert line["steps"] == 10 and line["warmup"] == 3
    # ranks other than 0 print nothing (torchrun launches the reference arm on every rank)
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(types.SimpleNamespace(gpus=2, steps=10, warmup=3))
    assert capsys.readouterr().out == ""
