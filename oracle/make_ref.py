"""TEST INFRASTRUCTURE ONLY -- recipe that stages the UNMODIFIED reference modules under oracle/_ref/.

    python oracle/make_ref.py            (run by __graft_entry__.build() whenever /root/reference exists)

The reference is pure Python: "building" it means copying the hot-path modules, byte for byte, from where they lie
under /root/reference into oracle/_ref/ (git-ignored, so no reference source enters the history; NOT gpurun-ignored,
so the files travel to the GPU box, where /root/reference does not exist).  oracle/ref_loader.py imports them from
there on top of oracle/shim (pure-torch stand-ins for torch_geometric / matplotlib).  Used only by
tests/, __graft_entry__.smoke() and bench.py's CPU / reference legs -- never by the product path.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SGS_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("model.py", "sampling.py", "utils.py", "training.py", "training_hybrid.py", "training_straight_through.py",
         "training_two_pass.py", "evaluate.py")


def stage(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "training_hybrid.py")):
        if verbose:
            print(f"make_ref: no reference at {SRC}; keeping whatever oracle/_ref already holds")
        return False
    os.makedirs(DST, exist_ok=True)
    lines = []
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        lines.append(f"{hashlib.sha256(open(os.path.join(DST, f), 'rb').read()).hexdigest()}  {f}")
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if verbose:
        print(f"make_ref: staged {len(FILES)} reference modules from {SRC} into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
