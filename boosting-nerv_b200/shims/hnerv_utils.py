"""Opt-in shim: `hnerv_utils` with the loss on the device kernels.

The reference's train loops do `from hnerv_utils import *` (train_nerv_all.py:21, train_nerv_compression.py:24) and
call `loss_fn(pred, target, args.loss)` (train_nerv_all.py:344).  Putting THIS directory ahead of the reference tree on
PYTHONPATH makes that import resolve here: the module executes the reference's own `hnerv_utils.py` (the next one on
sys.path) in its namespace - datasets, metrics, logging helpers, everything - and then replaces `loss_fn` by
`bnerv_b200.losses.loss_fn` (same signature and loss_type names; SSIM / MS-SSIM terms on csrc/loss_ops.cu) for CUDA
tensors.  CPU tensors keep the reference's own implementation.  `quant_tensor` (the post-training quantiser evaluate() and
quant_model() call, train_nerv_all.py:542,634) is routed to `bnerv_b200.ptq.quant_tensor` the same way (bit-identical).

    PYTHONPATH=/path/to/repo/boosting-nerv_b200/shims:/path/to/repo/boosting-nerv_b200:$PYTHONPATH python train_nerv_all.py ...
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_found = None
for _p in sys.path:
    _cand = os.path.join(_p or ".", "hnerv_utils.py")
    if os.path.isfile(_cand) and os.path.dirname(os.path.abspath(_cand)) != _here:
        _found = _cand
        break
if _found is None:
    raise ImportError("bnerv_b200 hnerv_utils shim: the reference's hnerv_utils.py is not on sys.path behind this directory")
_spec = importlib.util.spec_from_file_location("_reference_hnerv_utils", _found)
_ref = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_ref)
globals().update({k: v for k, v in vars(_ref).items() if not (k.startswith("__") and k.endswith("__"))})
if hasattr(_ref, "__all__"):
    __all__ = list(_ref.__all__)

_reference_loss_fn = _ref.loss_fn


def loss_fn(pred, target, loss_type="L2", batch_average=True):
    if pred.is_cuda:
        from bnerv_b200.losses import loss_fn as _native
        return _native(pred, target, loss_type, batch_average)
    return _reference_loss_fn(pred, target, loss_type, batch_average)


_reference_quant_tensor = _ref.quant_tensor      # hnerv_utils.py:101


def quant_tensor(t, bits=8):
    if t.is_cuda and t.dtype.is_floating_point and t.element_size() == 4 and 1 <= bits <= 8:
        from bnerv_b200.ptq import quant_tensor as _native
        return _native(t, bits)
    return _reference_quant_tensor(t, bits)
