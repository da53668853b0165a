"""Array plumbing between callers (numpy / torch, host / device) and the C ABI.

numpy or CPU-torch input  -> host entry points (skb_*_host), results returned as numpy / CPU torch.
CUDA torch input          -> device entry points on the caller's current stream, results are CUDA
                             tensors on the same device.
The compute dtype follows the input (float32 or float64; anything else is promoted to float64,
the reference's dtype) unless ``skmp_b200.config.compute_dtype`` forces one.
"""
import contextlib
import ctypes as C

import numpy as np

from . import _lib

try:  # torch is plumbing (device memory, streams, pinned host buffers), not the product
    import torch
except Exception:  # pragma: no cover
    torch = None


class config:
    compute_dtype = None          # None (follow input) | "float32" | "float64"
    pinned_threshold_bytes = 1 << 20


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def current_device_index() -> int:
    """Index of the CUDA device a native call issued now would run on (0 without torch / without a device: the call then
    fails loudly inside the library)."""
    if torch is not None and torch.cuda.is_available():
        return int(torch.cuda.current_device())
    return 0


def device_guard(index: int):
    """Context manager making cuda:``index`` current (native handles are created on, and bound to, one device)."""
    if torch is not None and torch.cuda.is_available():
        return torch.cuda.device(index)
    return contextlib.nullcontext()


class Batch:
    """Normalises one (N, D) batch of configurations and allocates matching outputs."""

    def __init__(self, qs, dim_cspace: int):
        forced = config.compute_dtype
        self.is_torch = _is_torch(qs)
        self.on_device = self.is_torch and qs.is_cuda
        if self.is_torch:
            if qs.dim() != 2:
                raise ValueError("qs must be 2-dimensional (n_point, dim_cspace)")
            dt = qs.dtype if qs.dtype in (torch.float32, torch.float64) else torch.float64
            if forced is not None:
                dt = getattr(torch, forced)
            self.q = qs.detach().to(dt).contiguous()
            self.np_dtype = np.float32 if dt == torch.float32 else np.float64
            self.torch_dtype = dt
            self.device = qs.device
        else:
            qs = np.asarray(qs)
            if qs.ndim != 2:
                raise ValueError("qs must be 2-dimensional (n_point, dim_cspace)")
            dt = qs.dtype if qs.dtype in (np.float32, np.float64) else np.float64
            if forced is not None:
                dt = np.dtype(forced)
            self.q = np.ascontiguousarray(qs, dtype=dt)
            self.np_dtype = np.dtype(dt).type
            self.torch_dtype = None
            self.device = None
        self.n = int(self.q.shape[0])
        if int(self.q.shape[1]) != dim_cspace:
            raise ValueError(f"qs has {self.q.shape[1]} columns, the kinematics map has dim_cspace={dim_cspace}")
        self.dtype_code = _lib.F32 if self.np_dtype == np.float32 else _lib.F64
        if self.on_device:
            _lib.require_cuda()
            self.stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self.device_index = int(self.device.index if self.device.index is not None else torch.cuda.current_device())
        else:
            self.stream = None
            self.device_index = current_device_index()       # host path: the current device does the work

    def device_ctx(self):
        """Make the tensor's device current for the duration of a native call."""
        if self.on_device:
            return torch.cuda.device(self.device)
        return contextlib.nullcontext()

    @property
    def q_ptr(self):
        if self.is_torch:
            return C.c_void_p(self.q.data_ptr())
        return C.c_void_p(self.q.ctypes.data)

    def empty(self, shape, dtype=None):
        if self.on_device:
            return torch.empty(shape, dtype=dtype or self.torch_dtype, device=self.device)
        np_dt = dtype or self.np_dtype
        nbytes = int(np.prod(shape)) * np.dtype(np_dt).itemsize
        if torch is not None and nbytes >= config.pinned_threshold_bytes and torch.cuda.is_available():
            tdt = {np.float32: torch.float32, np.float64: torch.float64, np.uint8: torch.uint8}[np.dtype(np_dt).type]
            return torch.empty(shape, dtype=tdt, pin_memory=True)
        return np.empty(shape, dtype=np_dt)

    def empty_bytes(self, n):
        if self.on_device:
            return torch.empty((n,), dtype=torch.uint8, device=self.device)
        return np.empty((n,), dtype=np.uint8)

    @staticmethod
    def ptr(arr):
        if arr is None:
            return None
        if _is_torch(arr):
            return C.c_void_p(arr.data_ptr())
        return C.c_void_p(arr.ctypes.data)

    def out(self, arr):
        """Convert an output buffer to the caller's array kind."""
        if arr is None:
            return None
        if self.is_torch:
            return arr if _is_torch(arr) else torch.from_numpy(arr)
        return arr.numpy() if _is_torch(arr) else arr

    def empty0(self):
        if self.is_torch:
            return torch.empty(0, dtype=self.torch_dtype, device=self.device)
        return np.empty(0)

    def nan11(self):
        """``AbstractConst.dummy_jacobian`` (skmp/constraint.py:55-56)."""
        if self.is_torch:
            return torch.full((1, 1), float("nan"), dtype=self.torch_dtype, device=self.device)
        return np.array([[np.nan]])
