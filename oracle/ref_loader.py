"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the reference's own modules UNMODIFIED from /root/reference on top of the
shim in oracle/shim.  /root/reference exists only in
the builder container, so this loader is used for two things only:
  * tests/golden/make_golden.py  (generates committed fixtures), and
  * `-m "not gpu"` tests that validate oracle/extended.py against the reference
    when the reference is present (skipped otherwise).
Nothing that runs on the GPU box may call `load()`.
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_DIR = os.environ.get("SGS_REFERENCE_DIR", "/root/reference")
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
