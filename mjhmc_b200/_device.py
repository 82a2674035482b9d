"""Device plumbing: torch is used only for device memory, streams and dtype bookkeeping."""
import ctypes as C

import numpy as np
import torch

from . import _lib

_TORCH = {"float32": torch.float32, "float64": torch.float64}
_CODE = {"float32": _lib.F32, "float64": _lib.F64}


def norm_dtype(dtype):
    name = np.dtype(dtype).name if not isinstance(dtype, torch.dtype) else str(dtype).replace("torch.", "")
    if name not in _TORCH:
        raise ValueError("dtype must be float32 or float64, got %r" % (dtype,))
    return name


def torch_dtype(dtype):
    return _TORCH[norm_dtype(dtype)]


def dtype_code(dtype):
    return _CODE[norm_dtype(dtype)]


def require_cuda(device=None):
    """The product path runs on the GPU only; fail loudly otherwise."""
    if not torch.cuda.is_available():
        raise RuntimeError("mjhmc_b200 needs a CUDA device: the sampler loop has no CPU fallback")
    _lib.load()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def to_device(a, dtype, device):
    """numpy / torch array -> contiguous device tensor of `dtype` (a copy unless already right)."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch_dtype(dtype)).contiguous()
    arr = np.ascontiguousarray(a)
    if not arr.flags.writeable:          # e.g. arrays out of np.load: torch refuses read-only memory
        arr = arr.copy()
    return torch.as_tensor(arr, device=device).to(torch_dtype(dtype)).contiguous()
