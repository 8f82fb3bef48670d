"""Launcher that runs the reference's own entry script on top of the drop-in modules.

    python -m sgs_gnn_b200.launch /path/to/SGS-GNN/main.py --mode learned --pipeline hybrid --GNN GCN ...

Why a launcher: `python main.py` puts the script's own directory at sys.path[0], AHEAD of PYTHONPATH, so the
reference's model.py / sampling.py / training*.py / utils.py would win over any shadow directory and the run would
silently use the reference code.  Here the accelerated modules are installed in `sys.modules` under the reference's
module names BEFORE the script runs, so `import model`, `from sampling import *`, ... (main.py:17-27,
training_hybrid.py:1-2, evaluate.py:2-3) resolve to them whatever sys.path says.

Every overlay module starts from the reference's ORIGINAL module of the same name (loaded by file path from the
script's directory) when that imports cleanly, and overlays only the accelerated symbols; names this build does not
accelerate (plotting helpers in utils.py, get_dataset in datasets.py, ...) therefore keep working.  When the original
cannot be imported (missing optional dependency) the overlay is the accelerated module alone.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import runpy
import sys
import types
import warnings

# reference module name -> (accelerated module, names to overlay; None = every public name)
OVERLAYS = {
    "utils": ("sgs_gnn_b200.utils", ("calculate_f1", "consistency_loss", "fix_seeds", "GpuMemoryProfiler")),
    "sampling": ("sgs_gnn_b200.sampling", ("gumbel_softmax_sampling", "random_edge_sampling")),
    "model": ("sgs_gnn_b200.model", ("GNNModel", "EdgeProbGCN", "EdgeProbMLP", "EdgeProbSAGE", "get_edge_mlp")),
    "datasets": ("sgs_gnn_b200.datasets", ("add_degree",)),
    "training_hybrid": ("sgs_gnn_b200.training_hybrid", ("train",)),
    "training_straight_through": ("sgs_gnn_b200.training_straight_through", ("train",)),
    "training_two_pass": ("sgs_gnn_b200.training_two_pass", ("train",)),
    "training": ("sgs_gnn_b200.training", ("train",)),
    "evaluate": ("sgs_gnn_b200.evaluate", ("evaluate", "ensemble_evaluate")),
}
ORDER = ("sampling", "utils", "model", "datasets", "training_hybrid", "training_straight_through",
         "training_two_pass", "training", "evaluate")


def _load_original(name, ref_dir):
    path = os.path.join(ref_dir, name + ".py")
    if not os.path.isfile(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod      # so that the original's own `from utils import *` etc. see the overlays made so far
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
    except Exception as exc:     # optional dependency missing (matplotlib, ogb, ...): accelerated names only
        sys.modules.pop(name, None)
        print(f"[sgs_gnn_b200.launch] {name}.py of the reference not importable ({type(exc).__name__}: {exc}); "
              f"using the accelerated module alone", file=sys.stderr)
        return None
    return mod


def install(ref_dir):
    """Install the overlay modules for the reference checkout at `ref_dir`; returns {name: module}."""
    installed = {}
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)  # the non-overlaid modules of the reference (parser.py, DeviceDir.py, ...)
    for name in ORDER:
        fast_name, names = OVERLAYS[name]
        fast = importlib.import_module(fast_name)
        base = _load_original(name, ref_dir)
        if base is None:
            mod = types.ModuleType(name)
            mod.__file__ = getattr(fast, "__file__", None)
            for k in dir(fast):
                if not k.startswith("_"):
                    setattr(mod, k, getattr(fast, k))
        else:
            mod = base
        for k in names:
            setattr(mod, k, getattr(fast, k))
        mod.__sgs_b200_overlay__ = fast_name
        sys.modules[name] = mod
        installed[name] = mod
    return installed


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    ref_dir = os.path.dirname(script)
    from . import _lib
    _lib.lib()                       # fail loudly now if the CUDA extension is missing
    install(ref_dir)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
