"""ctypes binding of libupretinex_b200.so (C ABI: include/upretinex_b200.h) + tensor-level wrappers.

PyTorch is plumbing here (device memory, streams); every op is one call into the C ABI with raw
device pointers and the current CUDA stream.  Missing library or missing GPU -> exception; there
is deliberately no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional, Tuple

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libupretinex_b200.so")

_lib = None
_lock = threading.Lock()

STATUS = {0: "UPR_OK", -1: "UPR_E_NULL", -2: "UPR_E_SHAPE", -3: "UPR_E_WORKSPACE", -4: "UPR_E_PARAM",
          -5: "UPR_E_DEVICE"}


class UprError(RuntimeError):
    def __init__(self, status: int, what: str):
        self.status = status
        super().__init__(f"{what} failed: status {status} ({status_string(status)})")


def _declare(lib):
    vp, i32, f64, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    lib.upr_version.restype = C.c_char_p
    lib.upr_status_string.restype = C.c_char_p
    lib.upr_status_string.argtypes = [i32]
    lib.upr_device_check.restype = i32
    lib.upr_clahe_workspace_bytes.restype = sz
    lib.upr_clahe_workspace_bytes.argtypes = [i32] * 5
    lib.upr_clahe_lab_f32.restype = i32
    lib.upr_clahe_lab_f32.argtypes = [vp, vp, i32, i32, i32, f64, i32, i32, vp, sz, vp]
    lib.upr_clahe_debug_dump.restype = i32
    lib.upr_clahe_debug_dump.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.upr_get_tables.restype = i32
    lib.upr_get_tables.argtypes = [vp] * 4


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  upretinex-b200 has no CPU fallback.")
                handle = C.CDLL(LIB_PATH)
                _declare(handle)
                _lib = handle
    return _lib


def status_string(status: int) -> str:
    if status <= 0:
        return STATUS.get(status, "UPR_E_?")
    try:
        return lib().upr_status_string(status).decode()
    except Exception:  # pragma: no cover
        return f"cudaError {status}"


def check(status: int, what: str) -> None:
    if status != 0:
        raise UprError(status, what)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (upretinex-b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


# one growing workspace per (device, stream): the C ABI never allocates
_workspaces = {}


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def release_workspaces() -> None:
    _workspaces.clear()


# ------------------------------------------------------------------------------------------------
# a1: CLAHE in Lab
# ------------------------------------------------------------------------------------------------
def clahe_lab(x: torch.Tensor, clip_limit: float = 2.0, tiles: Tuple[int, int] = (8, 8),
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [N,3,H,W] f32 CUDA -> [N,3,H,W] f32 CUDA (upr_clahe_lab_f32)."""
    x = _require_cuda_f32(x, "x")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"expected [N,3,H,W], got {tuple(x.shape)}")
    n, _, h, w = x.shape
    tx, ty = int(tiles[0]), int(tiles[1])
    if out is None:
        out = torch.empty_like(x)
    elif out.shape != x.shape or not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 CUDA tensor shaped like x")
    L = lib()
    with torch.cuda.device(x.device):
        nbytes = L.upr_clahe_workspace_bytes(n, h, w, tx, ty)
        if nbytes == 0:
            raise UprError(-2, "upr_clahe_workspace_bytes")
        ws = workspace(nbytes, x.device)
        check(L.upr_clahe_lab_f32(x.data_ptr(), out.data_ptr(), n, h, w, float(clip_limit), tx, ty,
                                  ws.data_ptr(), ws.numel(), _stream()), "upr_clahe_lab_f32")
    return out


def clahe_debug(x_shape, tiles: Tuple[int, int] = (8, 8), device=None, want_lab: bool = True):
    """Histograms / LUTs / Lab intermediate of the last clahe_lab() call on this stream."""
    n, _, h, w = x_shape
    tx, ty = int(tiles[0]), int(tiles[1])
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L = lib()
    with torch.cuda.device(device):
        ws = workspace(L.upr_clahe_workspace_bytes(n, h, w, tx, ty), device)
        hist = torch.empty((n, ty * tx, 256), dtype=torch.int32, device=device)
        lut = torch.empty((n, ty * tx, 256), dtype=torch.uint8, device=device)
        lab = torch.empty((n, 3, h, w), dtype=torch.uint8, device=device) if want_lab else None
        check(L.upr_clahe_debug_dump(ws.data_ptr(), n, h, w, tx, ty, hist.data_ptr(), lut.data_ptr(),
                                     lab.data_ptr() if lab is not None else None, _stream()), "upr_clahe_debug_dump")
    return hist, lut, lab


def tables():
    import numpy as np
    g = np.zeros(256, np.uint16); c = np.zeros(2048, np.uint16)
    yf = np.zeros(256, np.uint32); ig = np.zeros(4096, np.uint8)
    check(lib().upr_get_tables(g.ctypes.data, c.ctypes.data, yf.ctypes.data, ig.ctypes.data), "upr_get_tables")
    return {"gamma": g, "cbrt": c, "ify": (yf & 0xFFFF).astype(np.uint16), "y": (yf >> 16).astype(np.uint16),
            "invgamma": ig}
