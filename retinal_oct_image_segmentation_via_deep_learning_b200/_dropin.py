"""Glue between the reference-shaped scalar API ``f(y_true, y_pred)`` and the batched CUDA path.

The reference functions take two host arrays holding a binary mask and return one numpy scalar.
Here the arrays are validated, shipped to the GPU as uint8, reduced by the kernels of
``liboctm.so`` to exact integers, and the scalar is formed on the host with the reference's
expression (``derive.py``).  CUDA tensors are accepted as well and skip the upload.  There is no
CPU computation path: without a GPU or without the built library every function raises.
"""
from __future__ import annotations

import numpy as np
import torch

from . import suite


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("a CUDA device is required: this package has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_mask_u8(a, name="mask"):
    """numpy array / torch tensor holding a binary mask -> CUDA uint8 tensor of the same shape."""
    if isinstance(a, torch.Tensor):
        if a.dtype == torch.bool:
            t = a.view(torch.uint8) if a.is_contiguous() else a.to(torch.uint8)
        else:
            t = a.to(torch.uint8)
            if not bool((t == a).all()) or (t.numel() and int(t.max()) > 1):
                raise ValueError(f"{name}: the metric functions take binary masks (values 0/1)")
        return t.to(_device(), non_blocking=True).contiguous()
    arr = np.asarray(a)
    if arr.dtype == np.bool_:
        a8 = arr.view(np.uint8)
    else:
        a8 = arr.astype(np.uint8)
        if arr.size and (not np.array_equal(a8, arr) or a8.max() > 1):
            raise ValueError(f"{name}: the metric functions take binary masks (values 0/1)")
    return torch.from_numpy(np.ascontiguousarray(a8)).to(_device(), non_blocking=True)


def binary_counts(y_true, y_pred):
    """TP, FP, FN, TN, N of one binary mask pair as numpy int64 scalars (one K=2 confusion launch)."""
    t, p = as_mask_u8(y_true, "y_true"), as_mask_u8(y_pred, "y_pred")
    if t.shape != p.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {tuple(t.shape)} {tuple(p.shape)}")
    n = t.numel()
    if n == 0:
        z = np.int64(0)
        return z, z, z, z, z
    cm = suite.confusion(t.reshape(1, -1), p.reshape(1, -1), 2)[0].cpu().numpy()
    return np.int64(cm[1, 1]), np.int64(cm[0, 1]), np.int64(cm[1, 0]), np.int64(cm[0, 0]), np.int64(n)


def is_binary_like(a):
    """True when ``a`` really is a 0/1 mask: bool dtype, or uint8 whose maximum is <= 1 (checked).  Only such
    pairs may take the confusion-kernel shortcut of mean_squared_error / root_mean_squared_error / mad; any
    other integer array (multi-class label maps, uint8 images, boundary rows stored as uint8) goes through
    ``error_sums``, which reproduces the reference's ``astype(float)`` differences for any values."""
    if isinstance(a, torch.Tensor):
        if a.dtype == torch.bool:
            return True
        return a.dtype == torch.uint8 and (a.numel() == 0 or int(a.max()) <= 1)
    arr = np.asarray(a)
    if arr.dtype == np.bool_:
        return True
    return arr.dtype == np.uint8 and (arr.size == 0 or int(arr.max()) <= 1)


def error_sums(y_true, y_pred):
    """(sum d^2, sum |d|, n) for two arrays of any numeric dtype via the K3 kernels: exact int64 sums for
    integer-valued data (masks, boundary indices), float64 sums for floating data (soft positions)."""
    def to_t(a):
        if isinstance(a, torch.Tensor):
            return a
        arr = np.asarray(a)
        if arr.dtype.kind not in "iubf":
            raise TypeError(f"unsupported dtype {arr.dtype}")
        return torch.from_numpy(np.ascontiguousarray(arr))

    t, p = to_t(y_true), to_t(y_pred)
    if t.shape != p.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {tuple(t.shape)} {tuple(p.shape)}")
    n = t.numel()
    if n == 0:
        return np.int64(0), np.int64(0), 0
    floating = t.is_floating_point() or p.is_floating_point()
    if not floating:
        for a in (t, p):
            if a.dtype in (torch.int64, torch.uint64) and a.numel() and (int(a.max()) > 2**31 - 1 or int(a.min()) < -2**31):
                floating = True                       # does not fit the int32 kernel: take the float64 route
    dev = _device()
    if floating:
        t, p = t.to(torch.float64) if not t.is_floating_point() else t, p.to(torch.float64) if not p.is_floating_point() else p
        sq, ab = suite.boundary_error(t.to(dev).reshape(1, 1, -1), p.to(dev).reshape(1, 1, -1))
        return np.float64(sq.item()), np.float64(ab.item()), n
    t = (t.view(torch.uint8) if t.dtype == torch.bool else t).to(torch.int32)
    p = (p.view(torch.uint8) if p.dtype == torch.bool else p).to(torch.int32)
    sq, ab = suite.boundary_error(t.to(dev).reshape(1, 1, -1), p.to(dev).reshape(1, 1, -1))
    return np.int64(sq.item()), np.int64(ab.item()), n


integer_error_sums = error_sums          # former name


def contour_scalars(y_true, y_pred):
    """dict with hausdorff_distance / hausdorff_distance_95 / assd for one 2-D binary mask pair.

    Raises like the reference: ValueError for non-2-D or < 2x2 input (from find_contours),
    IndexError when a mask has no contour (``find_contours(...)[0]`` on an empty list)."""
    for a in (y_true, y_pred):
        shp = tuple(a.shape)
        if len(shp) != 2:
            raise ValueError("Only 2D arrays are supported.")
        if shp[0] < 2 or shp[1] < 2:
            raise ValueError("Input array must be at least 2x2.")
    t, p = as_mask_u8(y_true, "y_true"), as_mask_u8(y_pred, "y_pred")
    if t.shape != p.shape:
        # the reference computes each contour separately and would accept this; the batched kernels
        # need one shape
        raise ValueError("y_true and y_pred must have the same shape")
    ct = suite.contour_pass(t[None], p[None], 2)
    from . import derive
    m = derive.contour_metrics(ct.n_pts.cpu().numpy().view(np.uint32)[0, 1], ct.max_sq.cpu().numpy().view(np.uint32)[0, 1],
                               ct.p95_sq.cpu().numpy().view(np.uint32)[0, 1], ct.sum_dist.cpu().numpy()[0, 1])
    if not bool(m["contour_valid"]):
        raise IndexError("list index out of range")
    return m
