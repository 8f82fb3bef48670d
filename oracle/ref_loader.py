"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the reference's own modules UNMODIFIED on top of the shim in oracle/shim (SURVEY.md
section 8c / A.8): from /root/reference where it exists (the builder container), otherwise from
oracle/_ref/, the byte-for-byte staged copy oracle/make_ref.py puts there (git-ignored; it travels
to the GPU box with the snapshot, /root/reference does not).  Used by
  * tests/golden/make_golden.py  (generates committed fixtures),
  * tests that validate oracle/extended.py against the reference, and the `-m gpu` test that runs
    the reference's own training loop as the CALLER of the drop-in modules,
  * bench.py --impl reference (times the reference's own CPU implementation, kind "reference").
Never imported by the product path.
"""
import importlib
import os
import sys
import types
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_DIR = os.path.join(_HERE, "_ref")
REFERENCE_DIR = os.environ.get("SGS_REFERENCE_DIR", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_DIR, "training_hybrid.py")):
    REFERENCE_DIR = STAGED_DIR
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")

_REF_MODULES = ("sampling", "utils", "model", "training_hybrid",
                "training_straight_through", "training_two_pass", "training", "evaluate")
_SHIM_ROOTS = ("torch_geometric", "matplotlib")
_cache = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "training_hybrid.py"))


def load():
    """Return a namespace with the reference modules as attributes."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_DIR}")
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if k in _REF_MODULES or k.split(".")[0] in _SHIM_ROOTS}
    for k in saved_mods:
        del sys.modules[k]
    sys.path[:0] = [SHIM_DIR, REFERENCE_DIR]
    ns = types.SimpleNamespace()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for name in _REF_MODULES:
                setattr(ns, name, importlib.import_module(name))
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k in _REF_MODULES or k.split(".")[0] in _SHIM_ROOTS:
                mod = sys.modules.pop(k)
                if k.split(".")[0] in _SHIM_ROOTS:
                    setattr(ns, "_" + k.replace(".", "_"), mod)
        sys.modules.update(saved_mods)
    _cache = ns
    return ns


def load_callers(dropin_dir, names=("training_hybrid", "training_straight_through", "training_two_pass", "evaluate")):
    """The reference's own training / evaluation loops (unmodified files) bound to the DROP-IN modules: while they
    are imported, `sampling`, `utils` and `model` resolve to <dropin_dir>/{sampling,utils,model}.py, exactly what
    happens under `python -m sgs_gnn_b200.launch <reference>/main.py`.  Returns a namespace of the loaded modules;
    sys.modules / sys.path are restored afterwards."""
    import importlib.util
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_DIR}")
    shadow = ("sampling", "utils", "model")
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules[k] for k in list(sys.modules) if k in shadow or k in names}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, dropin_dir)
    ns = types.SimpleNamespace()
    try:
        for k in shadow:
            setattr(ns, k, importlib.import_module(k))       # the drop-in modules
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for name in names:
                spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_DIR, name + ".py"))
                mod = importlib.util.module_from_spec(spec)
                sys.modules[name] = mod
                spec.loader.exec_module(mod)
                setattr(ns, name, mod)
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k in shadow or k in names:
                del sys.modules[k]
        sys.modules.update(saved_mods)
    return ns
