"""The C-ABI library loads and exports every symbol include/qst.h declares (no compute calls:
this runs without a GPU).  Also pins the struct layouts shared between C and ctypes."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()          # nvcc cross-compiles for sm_100a; no GPU needed
    import qst_b200
    return qst_b200._lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "qst.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qst_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    import qst_b200
    names = _declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"libqst.so does not export {name}"
        assert name in qst_b200._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(qst_b200._lib.SIGNATURES) == names


def _declared_prototypes():
    """(return type, name, [parameter types]) of every prototype in include/qst.h, names stripped."""
    text = open(os.path.join(ROOT, "include", "qst.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = []
    for ret, name, args in re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(qst_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        params = []
        for a in args.split(","):
            a = re.sub(r"\bconst\b", "", a).strip()
            if not a or a == "void":
                continue
            if "*" in a:
                params.append("ptr")
            else:
                words = a.split()
                params.append(" ".join(words[:-1] if len(words) > 1 else words))
        out.append((re.sub(r"\bconst\b", "", ret).strip(), name, params))
    return out


def test_ctypes_signatures_agree_with_the_header():
    """Parameter count and width of every ctypes binding against the prototype in include/qst.h: a
    64-bit size bound as a C int, or one argument too few, corrupts a call without any error."""
    import qst_b200
    scalar = {"int": C.c_int, "int64_t": C.c_int64, "size_t": C.c_size_t, "float": C.c_float,
              "uint32_t": C.c_uint32, "long long": C.c_longlong}
    protos = _declared_prototypes()
    assert sorted(n for _, n, _ in protos) == sorted(qst_b200._lib.SIGNATURES)

    def is_pointer(t):
        return t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents")      # POINTER(...) types have .contents

    for ret, name, params in protos:
        restype, argtypes = qst_b200._lib.SIGNATURES[name]
        assert len(params) == len(argtypes), f"{name}: {len(params)} parameters declared, {len(argtypes)} bound"
        if "*" in ret:
            assert is_pointer(restype), name
        else:
            assert restype is scalar[ret], f"{name}: returns {ret}, bound as {restype}"
        for i, (want, got) in enumerate(zip(params, argtypes)):
            if want in ("ptr", "qst_stream_t"):
                assert is_pointer(got), f"{name} argument {i}: pointer declared, {got} bound"
            else:
                assert got is scalar[want], f"{name} argument {i}: {want} declared, {got} bound"


def test_version_and_error_channel(lib):
    assert lib.qst_version() == 100
    assert lib.qst_padded_dim(384) == 384 and lib.qst_padded_dim(385) == 448 and lib.qst_padded_dim(1) == 64
    import qst_b200
    plan = qst_b200._lib.TopkPlan()
    # invalid arguments are reported through the return code + qst_last_error, never a crash
    rc = lib.qst_topk_plan_make(0, 10, 8, 5, 0, 0, 148, C.byref(plan))
    assert rc == -1 and b"bad shape" in lib.qst_last_error()
    rc = lib.qst_topk_plan_make(10, 10, 8, 5000, 0, 0, 148, C.byref(plan))
    assert rc == -1 and b"k must be" in lib.qst_last_error()
    rc = lib.qst_topk_plan_make(10, 10, 8, 5, 0, 3, 148, C.byref(plan))      # unknown score function
    assert rc == -1
    assert lib.qst_topk_plan_make(10, 10, 8, 5, 0, 2, 148, C.byref(plan)) == 0 and plan.D_pad == 64
    assert lib.qst_padded_dim_for(768, 2) == 832 and lib.qst_padded_dim_for(768, 1) == 768


def test_compute_entries_reject_bad_arguments_before_any_device_work(lib):
    """No kernel is launched here (there is no GPU in this tier): every entry validates its arguments
    first and reports through the return code + qst_last_error instead of dereferencing a null pointer."""
    import qst_b200
    L = qst_b200._lib
    prm = L.QuadParams(0.6, 0.4, 1.0, 0.5, 0.5, 2.0, 1e-6, 0)
    bad_p = L.QuadParams(0.6, 0.4, 1.0, 0.5, 0.5, -1.0, 1e-6, 0)
    N = None
    cases = [
        (lambda: lib.qst_quadruplet_fwd(N, N, N, N, 7, 4, 8, C.byref(prm), 2, N, N, N, N), b"bad dtype 7"),
        (lambda: lib.qst_quadruplet_fwd(N, N, N, N, 0, -1, 8, C.byref(prm), 2, N, N, N, N), b"bad shape"),
        (lambda: lib.qst_quadruplet_fwd(N, N, N, N, 0, 4, 8, N, 2, N, N, N, N), b"null params"),
        (lambda: lib.qst_quadruplet_fwd(N, N, N, N, 0, 4, 8, C.byref(prm), 2, N, N, N, N), b"null loss_out"),
        (lambda: lib.qst_quadruplet_fwd(N, N, N, N, 0, 4, 8, C.byref(bad_p), 2, N, N, N, N), b"p must be positive"),
        (lambda: lib.qst_quadruplet_fwd(N, N, N, N, 0, 4, 8, C.byref(prm), 5, N, N, N, N), b"bad reduction"),
        (lambda: lib.qst_prep_rows(N, 0, 4, 8, 0, N, N, N, N, N, N), b"prep_rows"),
        (lambda: lib.qst_score_select(N, N, N, N, N), b"score_select"),
        (lambda: lib.qst_finalize_topk(N, N, N, N, N, N, N, N, 0, N, N, N, N), b"finalize_topk"),
        (lambda: lib.qst_exact_rescan(4, 10, 8, 3, 0, N, N, N, N, 0, N, N, N, N, N), b"exact_rescan"),
        (lambda: lib.qst_merge_topk(N, N, 2, 4, 10, N, N, N), b"merge_topk"),
        (lambda: lib.qst_ir_metrics(N, 4, 10, N, N, N, 0, N, N, N, N), b"ir_metrics"),
        (lambda: lib.qst_dense_scores(0, 10, 8, 0, N, N, N, N, N, N), b"dense_scores"),
        (lambda: lib.qst_quadruplet_eval(N, N, N, N, 0, 4, 8, N, N, N), b"quadruplet_eval"),
        (lambda: lib.qst_peer_barrier(N, 1, N), b"peer_barrier"),
        (lambda: lib.qst_exchange_candidates(N, N, N, 4, 3, N), b"null argument"),
    ]
    for call, message in cases:
        assert call() == -1
        assert message in lib.qst_last_error(), (message, lib.qst_last_error())


def test_plan_invariants(lib):
    """Host-side planning (pure arithmetic, no device): tiling, stripes, workspace layout."""
    import qst_b200
    for ctas in ("2", "1"):
        os.environ["QST_SCORE_CTAS"] = ctas
        try:
            for Q, N, D, k in [(10_000, 1_000_000, 768, 100), (1000, 10_000, 384, 10), (1, 1, 8, 1),
                               (100_000, 1_250_000, 768, 100), (20, 9000, 64, 900), (300, 257, 100, 7)]:
                plan = qst_b200._lib.TopkPlan()
                assert lib.qst_topk_plan_make(Q, N, D, k, 0, 0, 148, C.byref(plan)) == 0
                assert plan.ctas == int(ctas) and plan.rows_per_unit == 128 * plan.ctas
                # query-stationary tiles: CTA pairs whose query block fits TMEM + 64 KB of smem
                assert plan.qs == (1 if plan.ctas == 2 and plan.D_pad <= 768 else 0)
                assert plan.D_pad % 64 == 0 and plan.D_pad >= D
                assert plan.kprime >= k and plan.kprime % 16 == 0
                assert min(16, plan.kprime) <= plan.kunit <= plan.kprime and plan.cap >= plan.kunit + 64
                assert plan.m_tiles * plan.rows_per_unit >= Q > (plan.m_tiles - 1) * plan.rows_per_unit
                assert plan.n_tiles * 256 >= N > (plan.n_tiles - 1) * 256
                assert plan.stripes * plan.tiles_per_stripe >= plan.n_tiles
                assert (plan.stripes - 1) * plan.tiles_per_stripe < plan.n_tiles      # no empty stripe
                assert plan.units == plan.m_tiles * plan.stripes
                assert 1 <= plan.grid <= 148 // plan.ctas
                assert plan.off_thr < plan.off_cnt < plan.off_uthr < plan.off_cand < plan.ws_bytes
                assert plan.ws_bytes - plan.off_cand == plan.units * plan.rows_per_unit * plan.cap * 8
        finally:
            os.environ.pop("QST_SCORE_CTAS", None)
    # config 3 fills the machine: at least 95 % of the CTA slots of the last wave are busy
    plan = qst_b200._lib.TopkPlan()
    lib.qst_topk_plan_make(10_000, 1_000_000, 768, 100, 0, 0, 148, C.byref(plan))
    groups = 148 // plan.ctas
    waves = -(-plan.units // groups)
    assert plan.units / (waves * groups) > 0.95


def test_plan_invariants_hold_on_a_random_sweep(lib):
    """5000 seeded random (Q, N, D, k, k', score, SM count) shapes incl. the awkward ones (1, 255/256/257,
    k = 900, D = 1 / 769 / 4096, 1-16 SMs): every accepted plan tiles the problem exactly, has no empty
    stripe, a workspace that matches its layout, and enough per-unit entries to deliver k' candidates;
    every rejected one is rejected through the error channel."""
    import random

    import qst_b200
    os.environ.pop("QST_SCORE_CTAS", None)
    rng = random.Random(1)
    accepted = 0
    for _ in range(5000):
        Q = rng.choice([1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 512, 1000, rng.randint(1, 5000),
                        rng.randint(1, 200_000)])
        N = rng.choice([1, 2, 255, 256, 257, 1000, rng.randint(1, 100_000), rng.randint(1, 20_000_000)])
        D = rng.choice([1, 3, 8, 63, 64, 65, 100, 384, 768, 769, 832, 1024, rng.randint(1, 4096)])
        k = rng.choice([1, 5, 10, 100, 900, rng.randint(1, 1024)])
        kp = rng.choice([0, 0, 0, k, k + 16, rng.randint(1, 2048)])
        score, sms = rng.choice([0, 1, 2]), rng.choice([148, 148, 148, 132, 16, 2, 1])
        p = qst_b200._lib.TopkPlan()
        rc = lib.qst_topk_plan_make(Q, N, D, k, kp, score, sms, C.byref(p))
        if rc != 0:
            assert rc == -1 and lib.qst_last_error(), (Q, N, D, k, kp)
            continue
        accepted += 1
        what = (Q, N, D, k, kp, score, sms)
        assert p.rows_per_unit == 128 * p.ctas and p.ctas in (1, 2), what
        assert p.D_pad % 64 == 0 and p.D_pad >= D, what
        assert p.kprime >= k and 1 <= p.kunit <= max(p.kprime, 16) and p.cap >= p.kunit, what
        assert p.m_tiles * p.rows_per_unit >= Q > (p.m_tiles - 1) * p.rows_per_unit, what
        assert p.n_tiles * 256 >= N > (p.n_tiles - 1) * 256, what
        assert p.stripes >= 1 and p.stripes * p.tiles_per_stripe >= p.n_tiles, what
        assert (p.stripes - 1) * p.tiles_per_stripe < p.n_tiles, what
        assert p.units == p.m_tiles * p.stripes and 1 <= p.grid <= max(1, sms // p.ctas), what
        assert p.off_thr < p.off_cnt < p.off_uthr < p.off_cand < p.ws_bytes, what
        assert p.ws_bytes - p.off_cand == p.units * p.rows_per_unit * p.cap * 8, what
        assert p.kunit * p.stripes >= min(p.kprime, N) or p.kunit >= min(p.kprime, N), what
    assert accepted > 3000


def test_plan_picks_tile_shape_and_stripes_for_small_batches(lib):
    """Without QST_SCORE_CTAS: single-CTA tiles (M=128) when the batch fits one tile or when its last
    256-row block would be at most half full (small batches only), CTA pairs otherwise; a small batch
    against a long corpus gets one stripe per CTA (up to 160)."""
    import qst_b200
    os.environ.pop("QST_SCORE_CTAS", None)
    want = {1: 1, 32: 1, 128: 1, 129: 2, 256: 2, 257: 1, 384: 1, 385: 2, 512: 2, 640: 1, 1000: 2, 2048: 2,
            2049: 2, 2100: 2, 10_000: 2, 100_000: 2}
    for Q, ctas in want.items():
        plan = qst_b200._lib.TopkPlan()
        assert lib.qst_topk_plan_make(Q, 1_000_000, 768, 100, 0, 0, 148, C.byref(plan)) == 0
        assert plan.ctas == ctas, (Q, plan.ctas, ctas)
        assert plan.stripes <= 160 and plan.grid <= 148 // plan.ctas
    plan = qst_b200._lib.TopkPlan()
    lib.qst_topk_plan_make(128, 1_000_000, 768, 100, 0, 0, 148, C.byref(plan))
    assert plan.units == plan.stripes >= 140 and plan.grid == plan.units      # one corpus pass, all SMs busy
    assert plan.kunit == 16 and plan.cap == 256


def test_header_is_c99_and_a_plain_c_program_links_and_runs(lib, tmp_path):
    """The boundary from the C side: ``tests/c_abi_probe.c`` is compiled as C99 (-pedantic: the header must
    not need C++), linked against the in-tree libqst.so alone, and run here -- host-only entries (version,
    planner, error channel, argument validation) need no device."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    libdir = os.path.join(ROOT, "quadruplet-sentence-transformer_b200")
    exe = str(tmp_path / "c_abi_probe")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(here, "c_abi_probe.c"), "-L", libdir, "-lqst", "-Wl,-rpath," + libdir, "-o", exe],
                   check=True, capture_output=True, text=True)
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "c_abi_probe ok" in run.stdout


def test_integration_md_binding_stub_loads_against_the_built_library(lib):
    """The ctypes stub INTEGRATION.md shows a maintainer (section B) is executed as written, with only the
    library path pointed at the in-tree build: it must bind real symbols with the header's arity."""
    import qst_b200
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if 'C.CDLL("libqst.so")' in b)
    ns = {}
    exec(compile(stub.replace('C.CDLL("libqst.so")', f"C.CDLL({qst_b200._lib.LIB_PATH!r})"), "INTEGRATION.md", "exec"), ns)
    assert C.sizeof(ns["QuadParams"]) == C.sizeof(qst_b200._lib.QuadParams) == 32
    assert [f[0] for f in ns["QuadParams"]._fields_] == [f[0] for f in qst_b200._lib.QuadParams._fields_]
    assert len(ns["_lib"].qst_quadruplet_fwd_bwd.argtypes) == len(qst_b200._lib.SIGNATURES["qst_quadruplet_fwd_bwd"][1])
    assert ns["_lib"].qst_quadruplet_workspace_bytes() == lib.qst_quadruplet_workspace_bytes() > 0
    assert callable(ns["fused_loss_and_grads"])


def test_struct_layouts():
    import qst_b200
    assert C.sizeof(qst_b200._lib.QuadParams) == 32
    assert qst_b200._lib.TopkPlan.ws_bytes.offset % 8 == 0
    assert C.sizeof(qst_b200._lib.TopkPlan) == 32 + 4 * 14 + 8 * 5 + 8     # + qs flag and padding
    assert qst_b200._lib.TopkPlan.qs.offset == 32 + 4 * 14 + 8 * 5
    assert C.sizeof(qst_b200._lib.Scatter) == 16 * 8 + 4 + 4 + 8 and qst_b200._lib.Scatter.rows_per_block.offset == 136


def test_product_refuses_cpu_tensors_and_has_no_oracle_dependency():
    """No CPU fallback: CPU tensors raise; the product package never imports oracle/."""
    import torch
    import qst_b200
    x = torch.zeros(2, 4)
    with pytest.raises(qst_b200.QstError):
        qst_b200.gamma_quadruplet_loss(x, x, x, x)
    with pytest.raises(qst_b200.QstError):
        qst_b200.prepare_rows(x, True)
    pkg = os.path.join(ROOT, "quadruplet-sentence-transformer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
