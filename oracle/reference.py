"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference (`/root/reference/minimagen`) as the parity oracle.

The reference is pure Python/PyTorch, so "the oracle" is the reference itself executed on the CPU in fp32, made
importable by two stand-in packages for its un-installable imports (oracle/shims: einops_exts, resize_right -- see the
headers there).  It exists only in the build container: nothing that runs on the GPU box may import this
module.  There, parity is checked against (i) the committed golden fixtures under tests/golden/ that
oracle/make_golden.py generated with this loader and (ii) oracle/restatement.py, a self-contained torch-CPU restatement
of the same path which tests/test_oracle.py pins against the real reference and the goldens.

Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may import anything under oracle/.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("MINIMAGEN_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "minimagen"))


def load():
    """Import and return the reference package (`minimagen`), or raise if it is not present."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    mod = sys.modules.get("minimagen")
    if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT) \
            and REFERENCE_ROOT not in "".join(getattr(mod, "__path__", [])):
        raise RuntimeError("`minimagen` is already aliased to another package (install_as_minimagen?)")
    ref = importlib.import_module("minimagen")
    for sub in ("helpers", "layers", "diffusion_model", "Unet", "Imagen"):
        importlib.import_module(f"minimagen.{sub}")
    return ref
