"""TEST INFRASTRUCTURE ONLY -- compatibility shims for importing the *unmodified* reference.

The reference (sitzikbs/DeepFit, mounted read-only at /root/reference in the authoring
container; it does NOT exist on the GPU box) targets torch 1.4 / numpy 1.x.  These
monkey-patches are applied *outside* the reference so that its hot-path modules import and
run on torch 2.11 / numpy 2.3:

  * ``torch.symeig``   removed            -> ``torch.linalg.eigh(UPLO='U' if upper else 'L')``
                       (used by models/DeepFit.py:471, tutorial/tutorial_utils.py:223)
  * ``torch.load``     weights_only=True  -> force ``weights_only=False`` (pickled Namespace,
                       tutorial/compute_normals.py:34)
  * ``np.int``         removed            -> ``int`` (dataset.py:287)

Nothing under ``deepfit_b200/`` may import this module (product code never touches oracle/).
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("DEEPFIT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "DeepFit.py"))


_applied = False


def apply_shims() -> None:
    global _applied
    if _applied:
        return
    import numpy as np
    import torch

    # torch >= 1.13 keeps a ``torch.symeig`` stub that only raises; replace it unconditionally.
    def _symeig(A, eigenvectors=False, upper=True):
        w, v = torch.linalg.eigh(A, UPLO="U" if upper else "L")
        return w, v
    torch.symeig = _symeig

    _orig_load = torch.load

    def _load(*args, **kwargs):
        kwargs.setdefault("weights_only", False)
        return _orig_load(*args, **kwargs)
    torch.load = _load

    if not hasattr(np, "int"):
        np.int = int
    _applied = True


def import_reference():
    """Import the reference's hot-path modules untouched.  Returns (DeepFit, tutorial_utils)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s (expected on the GPU box)" % REFERENCE_ROOT)
    apply_shims()
    for sub in ("utils", "models", "tutorial"):
        p = os.path.join(REFERENCE_ROOT, sub)
        if p not in sys.path:
            sys.path.insert(0, p)
    # The repo ships drop-in modules with the same names; make sure we get the reference's.
    saved = {}
    for name in ("DeepFit", "tutorial_utils", "normal_estimation_utils", "ThreeDmFVNet"):
        if name in sys.modules and REFERENCE_ROOT not in (getattr(sys.modules[name], "__file__", "") or ""):
            saved[name] = sys.modules.pop(name)
    import importlib
    ref_deepfit = importlib.import_module("DeepFit")
    ref_tu = importlib.import_module("tutorial_utils")
    assert REFERENCE_ROOT in ref_deepfit.__file__, ref_deepfit.__file__
    assert REFERENCE_ROOT in ref_tu.__file__, ref_tu.__file__
    # keep the reference modules under private names and restore whatever was there before
    sys.modules["_ref_DeepFit"] = ref_deepfit
    sys.modules["_ref_tutorial_utils"] = ref_tu
    for name in ("DeepFit", "tutorial_utils"):
        sys.modules.pop(name, None)
    sys.modules.update(saved)
    for sub in ("utils", "models", "tutorial"):
        p = os.path.join(REFERENCE_ROOT, sub)
        while p in sys.path:
            sys.path.remove(p)
    return ref_deepfit, ref_tu
