"""Imports the UNMODIFIED reference `model` package from /root/reference in THIS container only (the GPU box does not
have it). Two environment shims, no arithmetic change:
  1. the literal device 'cuda:1' (reference model/models.py:118-119, model/utils.py:72) is remapped to the run device;
  2. nothing else — on CPU `pack_padded_sequence` already gets CPU lengths.
Test infrastructure only: never imported by the product package."""
import contextlib
import os
import sys

import torch

REF_ROOT = os.environ.get("DUALVGR_REFERENCE", "/root/reference")
_RUN_DEVICE = "cpu"
_orig_to = torch.Tensor.to


def _patched_to(self, *args, **kwargs):
    args = list(args)
    if args and isinstance(args[0], str) and args[0] == "cuda:1":
        args[0] = _RUN_DEVICE
    if kwargs.get("device", None) == "cuda:1":
        kwargs["device"] = _RUN_DEVICE
    return _orig_to(self, *args, **kwargs)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


@contextlib.contextmanager
def reference_on_path():
    """Temporarily makes `import model.models` resolve to the reference (and evicts any cached `model*` modules)."""
    saved = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    torch.Tensor.to = _patched_to
    try:
        yield
    finally:
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def load_reference(device="cpu"):
    """Returns the reference's `model.models` module (kept alive by the caller) with the device shim active."""
    global _RUN_DEVICE
    _RUN_DEVICE = device
    import warnings
    warnings.filterwarnings("ignore")
    with reference_on_path():
        import model.models as modelset  # noqa
    torch.Tensor.to = _patched_to   # the shim must stay active while the reference modules run
    return modelset


def load_reference_losses():
    """The reference's top-level utils.py (common_loss, loss_dependence) loaded under a private module name.
    loss_dependence hard-codes .cuda() (utils.py:22); on CPU we patch Tensor.cuda to identity for the call."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_top_utils", os.path.join(REF_ROOT, "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
