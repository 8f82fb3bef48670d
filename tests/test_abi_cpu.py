"""CPU: the C-ABI library loads, exports every symbol include/sgs_b200.h declares with the
argument lists sgs_gnn_b200/_lib.py binds, and the host-side logic behaves (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from sgs_gnn_b200 import _lib, rng, synth


def _header_decls():
    src = open(os.path.join(ROOT, "include", "sgs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\n\s*(const char\*|int32_t|int64_t|size_t)\s+(sgs_\w+)\s*\(([^;]*?)\)\s*;", src)


def test_library_exports_every_declared_symbol():
    h = _lib.lib()
    decls = _header_decls()
    assert len(decls) >= 30
    for _ret, name, _args in decls:
        assert hasattr(h, name), f"{name} declared in sgs_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} not bound in _lib.py"
    assert set(_lib.SIGNATURES) == {d[1] for d in decls}
    assert h.sgs_version() == 100
    assert h.sgs_last_error() is not None


def test_ctypes_signatures_match_header():
    kinds = {"float": ctypes.c_float, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32,
             "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t, "sgs_stream_t": ctypes.c_void_p}
    for _ret, name, args in _header_decls():
        args = args.replace("\n", " ").strip()
        want = []
        if args not in ("void", ""):
            for a in args.split(","):
                a = a.strip()
                want.append(ctypes.c_void_p if "*" in a else kinds[a.replace("const ", "").split()[0]])
        got = _lib.SIGNATURES[name][1]
        assert len(got) == len(want), name
        for g, w in zip(got, want):
            assert ctypes.sizeof(g) == ctypes.sizeof(w) and (g is ctypes.c_float) == (w is ctypes.c_float), name


def test_workspace_queries_need_no_gpu():
    h = _lib.lib()
    assert h.sgs_csr_workspace_bytes(1000, 10) > 16000
    assert h.sgs_topq_workspace_bytes(1 << 20) >= 2 * 128 * 4
    assert h.sgs_edge_score_workspace_bytes(1000, 50, 256, 0, 0) >= 1000 * 3 * 256 * 4
    assert h.sgs_edge_score_workspace_bytes(1000, 50, 256, 0, 1) >= 1000 * 5 * 256 * 4


def test_bad_arguments_return_error_codes_not_crashes():
    h = _lib.lib()
    assert h.sgs_gemm(None, 1, 1, None, 1, 1, None, 4, 4, 4, 4, 0, 0, None) == -1
    assert b"null pointer" in h.sgs_last_error()
    assert h.sgs_topq_find(None, None, 1, 0, None) == -1
    assert h.sgs_spmm(None, None, None, None, None, None, None, 0, 4, None, None, 0, 0.0, 0, None) == -1


def test_ops_refuse_cpu_tensors():
    from sgs_gnn_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.sample_topq(torch.rand(16), torch.rand(16), 4)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.linear_nt(torch.rand(4, 4), torch.rand(4, 4))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsgs_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_rng_mirror_statistics():
    m = rng.keep_mask(1234, np.arange(2000), 256, 0.3)
    assert m.shape == (2000, 256)
    assert abs(m.mean() - 0.7) < 5e-3
    assert (rng.keep_mask(1234, [7], 256, 0.3) == m[7]).all()
    assert rng.keep_mask(1, [0], 8, 0.0).all()


def test_synth_graph_contract():
    b = synth.make_graph("smallcora", seed=7)
    n, e = b.num_nodes, b.num_edges
    assert (n, e, b.x.shape[1]) == (2708, 10556, 1433)
    ei = b.edge_index
    key = ei[0] * n + ei[1]
    assert bool((key[1:] > key[:-1]).all()) and bool((ei[0] != ei[1]).all())
    assert torch.equal(torch.sort(ei[1] * n + ei[0]).values, key)          # symmetric
    assert int(b.train_mask.sum()) == int(n * 0.2) and not bool((b.train_mask & b.val_mask).any())
    b2 = synth.make_graph("arxiv-year", seed=7, scale=0.01)
    assert b2.num_edges == round(1166243 * 0.01)


def test_dropin_modules_export_reference_names():
    import importlib
    import sys
    d = os.path.join(ROOT, "sgs_gnn_b200", "dropin")
    sys.path.insert(0, d)
    try:
        for mod, names in (("model", ["GNNModel", "EdgeProbGCN", "EdgeProbMLP", "get_edge_mlp"]),
                           ("sampling", ["gumbel_softmax_sampling", "random_edge_sampling"]),
                           ("utils", ["consistency_loss", "calculate_f1", "GpuMemoryProfiler", "fix_seeds"]),
                           ("training", ["train"]), ("training_hybrid", ["train"]),
                           ("training_straight_through", ["train"])):
            sys.modules.pop(mod, None)
            m = importlib.import_module(mod)
            for nme in names:
                assert hasattr(m, nme), (mod, nme)
            sys.modules.pop(mod, None)
    finally:
        sys.path.remove(d)


def test_launcher_makes_dropin_win_over_script_directory(tmp_path):
    """`python main.py` puts the script's directory first on sys.path, so a PYTHONPATH shadow never wins.
    The launcher installs the accelerated modules in sys.modules under the reference's names before the script runs:
    a stand-in main.py next to decoy model.py / utils.py / sampling.py must see the sgs_gnn_b200 classes, while names
    the build does not accelerate keep coming from the script's own modules."""
    import subprocess
    import sys
    (tmp_path / "model.py").write_text("class GNNModel:\n    origin = 'decoy'\nOTHER = 'kept'\n")
    (tmp_path / "utils.py").write_text("def calculate_f1(*a):\n    return 'decoy'\ndef plot_learning_curves():\n"
                                       "    return 'kept'\n")
    (tmp_path / "sampling.py").write_text("import this_module_does_not_exist\n")
    (tmp_path / "main.py").write_text(
        "import sys\nfrom model import *\nimport model, utils, sampling, training_hybrid\n"
        "from utils import calculate_f1, plot_learning_curves\n"
        "print('RESULT', GNNModel.__module__, model.OTHER, calculate_f1.__module__, plot_learning_curves(),\n"
        "      sampling.gumbel_softmax_sampling.__module__, training_hybrid.train.__module__, sys.argv[1:])\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "sgs_gnn_b200.launch", str(tmp_path / "main.py"), "--mode", "learned"],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0]
    assert line == ("RESULT sgs_gnn_b200.model kept sgs_gnn_b200.utils kept sgs_gnn_b200.sampling "
                    "sgs_gnn_b200.training_hybrid ['--mode', 'learned']"), line
    assert "sampling.py of the reference not importable" in r.stderr


def test_launcher_overlays_the_reference_checkout():
    """With the real reference modules present (/root/reference or oracle/_ref): the overlay keeps the reference's
    remaining names (utils.plot_learning_curves, visualize) and replaces the accelerated ones."""
    import sys
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference modules not staged")
    from sgs_gnn_b200 import launch
    saved = {k: sys.modules.get(k) for k in launch.ORDER}
    saved_path = list(sys.path)
    sys.path.insert(0, ref_loader.SHIM_DIR)      # stand-ins for torch_geometric / matplotlib (absent in this image)
    try:
        mods = launch.install(ref_loader.REFERENCE_DIR)
        assert mods["model"].GNNModel.__module__ == "sgs_gnn_b200.model"
        assert mods["utils"].calculate_f1.__module__ == "sgs_gnn_b200.utils"
        assert hasattr(mods["utils"], "plot_learning_curves") and hasattr(mods["utils"], "visualize")
        assert mods["training_hybrid"].train.__module__ == "sgs_gnn_b200.training_hybrid"
        assert mods["evaluate"].ensemble_evaluate.__module__ == "sgs_gnn_b200.evaluate"
    finally:
        sys.path[:] = saved_path
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in list(sys.modules):
            if k.split(".")[0] in ("torch_geometric", "matplotlib"):
                del sys.modules[k]
