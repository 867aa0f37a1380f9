"""Loader for the *live* reference (build container only).

TEST INFRASTRUCTURE -- never imported by the product package.

The reference tree (/root/reference, read-only) does not import on a modern
torch: ``QuantTorch/__init__.py:2`` pulls ``train`` -> ``optuna`` (absent) and
``layers/dorefa_layers.py:2`` / ``terner_layers.py:2`` import two names that
were removed from ``torch._jit_internal``.  This module installs a two-line
shim (no edits to the reference tree) and returns ``(functions, layers)``.

It is used only by ``oracle/gen_golden.py`` and by the CPU tests that pin the
oracle against the live reference when the tree is present.  ``/root/reference``
does not exist on the GPU box, so nothing on the ``-m gpu`` path calls this.
"""
import os
import sys
import types
import warnings


def reference_root():
    for cand in (os.environ.get("QT_REF_PATH"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "QuantTorch")):
            return cand
    return None


def load_reference():
    """Return (functions, layers) modules of the unmodified reference, or None."""
    root = reference_root()
    if root is None:
        return None
    import torch._jit_internal as ji
    if not hasattr(ji, "weak_module"):
        ji.weak_module = lambda c: c
    if not hasattr(ji, "weak_script_method"):
        ji.weak_script_method = lambda f: f
    if "QuantTorch" not in sys.modules:
        pkg = types.ModuleType("QuantTorch")
        pkg.__path__ = [os.path.join(root, "QuantTorch")]
        sys.modules["QuantTorch"] = pkg
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from QuantTorch import functions, layers  # noqa
    return functions, layers
