"""CPU-side checks: the C-ABI library loads and exports every symbol include/dgcn.h
declares, the drop-in modules keep the reference's constructor signatures and
state_dict keys, and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import golden_util as gu
from oracle import ref_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from deep_gcns_torch_b200 import _native, build
    build.build()
    header = open(os.path.join(ROOT, "include", "dgcn.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(dgcn_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 15
    handle = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    lib = _native.lib()
    assert lib.dgcn_version() >= 100
    assert lib.dgcn_status_string(-2).decode().startswith("request outside")
    # size queries are pure host code: callable without a GPU
    assert lib.dgcn_knn_graph_workspace_bytes(16, 64, 4096, 20) >= 16 * 4096 * 4
    assert lib.dgcn_knn_graph_workspace_bytes(16, 64, 4096, 540) >= 4096 * 4096 * 4
    assert lib.dgcn_csr_build_workspace_bytes(1000, 5000) > 3 * 5000 * 4


def test_struct_layouts_match_header():
    from deep_gcns_torch_b200 import _native
    assert ctypes.sizeof(_native.BasicConvC) == 96
    assert ctypes.sizeof(_native.DilationC) == 32
    assert ctypes.sizeof(_native.GenconvParamsC) == 80
    assert ctypes.sizeof(_native.CsrHubsC) == 40
    assert ctypes.sizeof(_native.GenconvFusionC) == 40
    assert ctypes.sizeof(_native.BlockFusionC) == 40


def test_no_cpu_fallback():
    from deep_gcns_torch_b200.gcn_lib import dense as D, sparse as S
    x = torch.randn(1, 4, 16, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        D.DynConv2d(4, 4, 3)(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        D.DenseDilatedKnnGraph(3)(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        S.GENConv(4, 4)(torch.randn(5, 4), torch.zeros((2, 3), dtype=torch.long))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "deep_gcns_torch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


@pytest.mark.parametrize("name", [n for n in gu.names("dense_") if "grid" not in n])
def test_dense_state_dict_compat(name):
    from deep_gcns_torch_b200.gcn_lib import dense as D
    c = gu.load(name)
    m = c.meta
    if "static" in name:
        mod = D.GraphConv2d(m["in_channels"], m["out_channels"], m["conv"], m["act"], m["norm"], m["bias"])
    else:
        mod = D.DynConv2d(m["in_channels"], m["out_channels"], m["k"], m["dilation"], m["conv"], m["act"],
                          m["norm"], m["bias"])
    own = mod.state_dict()
    assert list(own.keys()) == list(c.sd.keys())
    assert all(tuple(own[k].shape) == tuple(c.sd[k].shape) and own[k].dtype == c.sd[k].dtype for k in own)
    mod.load_state_dict(c.sd, strict=True)


@pytest.mark.parametrize("name", gu.names("sparse_"))
def test_sparse_state_dict_compat(name):
    from deep_gcns_torch_b200.gcn_lib import sparse as S
    c = gu.load(name)
    m = dict(c.meta)
    in_dim, emb_dim = m.pop("in_dim"), m.pop("emb_dim")
    m.pop("N")
    mod = S.GENConv(in_dim, emb_dim, **m)
    own = mod.state_dict()
    assert list(own.keys()) == list(c.sd.keys())
    assert all(tuple(own[k].shape) == tuple(c.sd[k].shape) for k in own)
    mod.load_state_dict(c.sd, strict=True)
    assert isinstance(getattr(mod, "t", 1.0), (float, torch.nn.Parameter))


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree absent (GPU box)")
def test_signatures_match_live_reference():
    import inspect
    from deep_gcns_torch_b200.gcn_lib import dense as D, sparse as S
    rd, rs = ref_shims.load_reference()
    pairs = [(getattr(D, n), getattr(rd, n)) for n in
             ("DenseDilated", "DenseDilatedKnnGraph", "DilatedKnnGraph", "MRConv2d", "EdgeConv2d", "GraphConv2d",
              "DynConv2d", "PlainDynBlock2d", "ResDynBlock2d", "DenseDynBlock2d", "BasicConv")]
    pairs += [(getattr(S, n), getattr(rs, n)) for n in ("GENConv", "MsgNorm", "MLP")]
    for mine, ref in pairs:
        a, b = inspect.signature(mine.__init__), inspect.signature(ref.__init__)
        assert [(p.name, p.default) for p in a.parameters.values()] == \
               [(p.name, p.default) for p in b.parameters.values()], mine.__name__
    # model stacks of the reference build on top of the drop-in classes unchanged
    import types
    opt = types.SimpleNamespace(n_filters=16, k=4, act="relu", norm="batch", bias=True, epsilon=0.2,
                                stochastic=True, conv="edge", n_blocks=3, block="res", in_channels=9,
                                n_classes=13, dropout=0.3)
    src = open(os.path.join(ref_shims.REFERENCE_ROOT, "examples/sem_seg_dense/architecture.py")).read()
    src = src.split('if __name__ == "__main__"')[0].replace("import __init__\n", "")
    src = src.replace("from gcn_lib.dense import", "from deep_gcns_torch_b200.gcn_lib.dense import")
    ns = {}
    exec(compile(src, "architecture.py", "exec"), ns)
    mine = ns["DenseDeepGCN"](opt)
    import sys
    sys.path.insert(0, os.path.join(ref_shims.REFERENCE_ROOT, "examples/sem_seg_dense"))
    ref_src = open(os.path.join(ref_shims.REFERENCE_ROOT, "examples/sem_seg_dense/architecture.py")).read()
    ref_src = ref_src.split('if __name__ == "__main__"')[0].replace("import __init__\n", "")
    ns2 = {"__name__": "ref_arch"}
    sys.path.insert(0, ref_shims.REFERENCE_ROOT)
    try:
        exec(compile(ref_src, "ref_architecture.py", "exec"), ns2)
    finally:
        sys.path.remove(ref_shims.REFERENCE_ROOT)
    theirs = ns2["DenseDeepGCN"](opt)
    assert list(mine.state_dict().keys()) == list(theirs.state_dict().keys())
    mine.load_state_dict(theirs.state_dict(), strict=True)


def test_bench_line_carries_every_contract_key():
    """Static check of bench.py's JSON line (no GPU): the dict literal assembled in run_native, evaluated
    with stand-in measurements, has every key of the bench contract and consistent derived values."""
    import ast
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "bench.py")
    tree = ast.parse(open(path).read())
    g = {"__file__": path, "__name__": "bench_static"}
    top = [n for n in tree.body if isinstance(n, (ast.Assign, ast.Import, ast.ImportFrom))]
    exec(compile(ast.Module(top, []), path, "exec"), g)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_native")
    lit = next(n.value for n in ast.walk(fn) if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Name)
               and n.targets[0].id == "out" and isinstance(n.value, ast.Dict))

    class Args:
        steps, warmup = 20, 1

    loc = dict(world=2, args=Args, ms=11.4, ms_e2e=12.7, clk={"sm_mhz": 1965.0}, achieved=220.0, peak=6563.9,
               traffic=72000000, peak_src="measured", kernel_ms=0.507, tensor_peak=1691.8, sm_max=1965.0,
               hbm_achieved=66.0, tensor_flop=1.1e11, numa={"numa_node": 0})
    out = eval(compile(ast.Expression(lit), path, "eval"), g, loc)
    json.dumps(out)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert key in out, key
    assert out["warmup"] >= 3 and out["n_gpus"] == 2 and out["scaling"] == "weak" and out["vs_baseline"] is None
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(out["roofline"])
    assert out["roofline"]["bound"] in ("hbm", "tensor") and out["roofline"]["unit"] == "TFLOP/s"
    assert abs(out["roofline"]["frac"] - 220.0 / 1691.8) < 1e-9 and out["roofline"]["hbm"]["unit"] == "GB/s"
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(out["e2e"])
    assert "workload" in out["config"] and "model" not in out["config"]
    edges = 16 * 4096 * 20
    assert abs(out["value"] - edges * 2 * 20 / 11.4e-3) < 1.0          # whole-job aggregate over both ranks
    assert out["gpu_launches"] > 0
