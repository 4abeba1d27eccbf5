"""N4: device-side preparation of the reference's output files.

The reference writes its stacks with np.save as (H, W, C) float64 arrays (scripts/2_feature_extraction.py:193-214) and its
label map as kmeans_result + 1 in a uint8 GeoTIFF (scripts/3_classification.py:394, extract.py:795-807).  Widening,
interleaving and the +1 / uint8 cast happen on the device; the host receives the exact file payload in page-locked memory and
only writes bytes (the GeoTIFF container itself stays with the reference's rasterio code).
"""
from __future__ import annotations

import io

import numpy as np
import torch

from . import _lib
from .device import ptr, require_cuda, stream_ptr


def npy_header(shape, dtype=np.float64) -> bytes:
    """The bytes np.save puts in front of the data for a C-ordered array of this shape (format 1.0 / 2.0 as numpy chooses)."""
    buf = io.BytesIO()
    np.lib.format.write_array_header_1_0(buf, {"descr": np.lib.format.dtype_to_descr(np.dtype(dtype)), "fortran_order": False,
                                               "shape": tuple(int(s) for s in shape)})
    return buf.getvalue()


def stack_to_npy_bytes(planes: torch.Tensor, n_px: int, H: int, W: int, n_channels: int = None) -> torch.Tensor:
    """Page-locked uint8 tensor holding the complete .npy file of the (H, W, C) float64 stack made of the first C planes."""
    require_cuda()
    C = int(n_channels if n_channels is not None else planes.shape[0])
    assert H * W == n_px and planes.dtype == torch.float32 and planes.is_contiguous()
    head = npy_header((H, W, C))
    assert len(head) % 8 == 0                                      # numpy pads the header so that the data is 64-byte aligned
    dev_payload = torch.empty(n_px * C, dtype=torch.float64, device=planes.device)
    _lib.call("rsx_planes_to_hwc_f64", ptr(planes), planes.stride(0), n_px, C, ptr(dev_payload), stream_ptr())
    out = torch.empty(len(head) + n_px * C * 8, dtype=torch.uint8, pin_memory=True)
    out[:len(head)] = torch.frombuffer(bytearray(head), dtype=torch.uint8)
    out[len(head):].view(torch.float64).copy_(dev_payload, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out


def save_stack_npy(path: str, planes: torch.Tensor, n_px: int, H: int, W: int, n_channels: int = None) -> None:
    """np.save(path, stack.astype(float64)) with the stack coming straight from the device."""
    data = stack_to_npy_bytes(planes, n_px, H, W, n_channels)
    with open(path if path.endswith(".npy") else path + ".npy", "wb") as f:
        f.write(memoryview(data.numpy()))


def labels_for_geotiff(labels_i32: torch.Tensor, H: int, W: int) -> np.ndarray:
    """(H, W) uint8 array = kmeans_result + 1, ready for the reference's save_classification_geotiff."""
    require_cuda()
    n = H * W
    dev = torch.empty(n, dtype=torch.uint8, device=labels_i32.device)
    _lib.call("rsx_labels_plus1_u8", ptr(labels_i32), n, ptr(dev), stream_ptr())
    out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    out.copy_(dev, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out.numpy().reshape(H, W)
