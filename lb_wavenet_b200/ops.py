"""Primitive ops of the reference (ops.py:4-39) backed by the CUDA library.

mu_encode / mu_decode run on the device through table-driven kernels that reproduce the
reference's float32 numpy twins bit for bit (n_quanta == 256).  Inputs may be numpy arrays or torch
tensors (host or device); the result is a device tensor.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _to_device(x, dtype):
    import torch
    if not torch.cuda.is_available():
        raise _lib.WaveNetLibError("lb_wavenet_b200.ops needs a CUDA device: there is no CPU fallback")
    t = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))
    return t.to(device="cuda", dtype=dtype).contiguous()


def mu_encode(x, n_quanta: int = 256):
    """mu-law encode and quantize (reference ops.py:4-9 / 23-28) -> int32 codes in [0, 255]."""
    import torch
    if n_quanta != 256:
        raise ValueError("only n_quanta == 256 is built")
    lib = _lib.load()
    xd = _to_device(x, torch.float32)
    q = torch.empty(xd.shape, dtype=torch.int32, device=xd.device)
    _lib.check(lib.wn_mu_encode(xd.data_ptr(), q.data_ptr(), xd.numel(), _lib.cur_stream()), "wn_mu_encode")
    return q


def mu_decode(quant, n_quanta: int = 256):
    """integer mu-law code -> pre-encoded float32 value (reference ops.py:12-20 / 31-39)."""
    import torch
    if n_quanta != 256:
        raise ValueError("only n_quanta == 256 is built")
    lib = _lib.load()
    qd = _to_device(quant, torch.int32)
    x = torch.empty(qd.shape, dtype=torch.float32, device=qd.device)
    _lib.check(lib.wn_mu_decode(qd.data_ptr(), x.data_ptr(), qd.numel(), _lib.cur_stream()), "wn_mu_decode")
    return x
