"""dualvgr-videoqa_b200 — B200-native (sm_100a) implementation of the DualVGR reasoning core.

Layout
  csrc/       hand-written CUDA kernels + the C ABI (include/dualvgr_b200.h) -> libdualvgr_b200.so
  _lib.py     ctypes binding of the C ABI (fails loudly when the library is missing: there is no CPU fallback)
  ops.py      torch.autograd.Function wrappers (device memory + streams are torch's, arithmetic is the library's)
  model/      drop-in mirror of the reference's ``model`` package (model.models.DualVGR, same state_dict keys)
  utils.py    mirror of the reference's top-level utils.py (todevice, common_loss, loss_dependence)
  engine.py   data-parallel train step (flat fp32 parameter/gradient buffers, NCCL allreduce, fused clip + Adam)
"""
__version__ = "0.1.0"
