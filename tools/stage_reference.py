"""Stage the UNMODIFIED reference checkout under baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box) so that
the drop-in boundary can be tested against the reference's own code on the B200 (tests/test_gpu_dropin.py) and the reference's PyTorch
modules can be timed eagerly on the same GPU (bench.py `torch_eager_b200`).  Only the Python the model path needs is staged (models/,
utils/, functions.py, main.py, configs/); nothing under baseline/_ref/ is ever committed or imported by the product package.

    python tools/stage_reference.py            # no-op when /root/reference is absent (the GPU box uses the staged copy)
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("EFFCONF_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
ITEMS = ("models", "utils", "functions.py", "main.py", "configs", "LICENSE")


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: keeping {DST} as it is ({'present' if os.path.isdir(DST) else 'absent'})")
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    for item in ITEMS:
        s, d = os.path.join(SRC, item), os.path.join(DST, item)
        if not os.path.exists(s):
            continue
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    if verbose:
        print("staged", ", ".join(ITEMS), "->", DST)
    return True


def import_reference():
    """Make the staged reference importable: stub the four uninstalled third-party modules the encoder / CTC path never calls
    (jiwer, ctcdecode, warp_rnnt, kenlm) and put baseline/_ref first on sys.path.  Returns the staged path or None."""
    import types
    if not os.path.isdir(os.path.join(DST, "models")):
        return None
    for n in ("jiwer", "ctcdecode", "warp_rnnt", "kenlm"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    if not hasattr(sys.modules["ctcdecode"], "CTCBeamDecoder"):
        sys.modules["ctcdecode"].CTCBeamDecoder = object
    if DST not in sys.path:
        sys.path.insert(0, DST)
    return DST


if __name__ == "__main__":
    stage()
