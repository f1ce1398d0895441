"""The reference's training metric on the device: Jaccard (deeplabv3p/metrics.py:30-45, compiled in at train.py:141).

The pixel counting (per image, per class: intersection, true, predicted) is a libdlv3p kernel (dlv3p_op_jaccard_counts, integer, bit exact);
what is left is arithmetic on 3*(NC+1) numbers per image, done here in float64 exactly in the reference's order:
    for i in 0..NC:  ious_b = inter_b / union_b ; iou_i = mean over the images that contain class i (NaN if none)
    Jaccard = mean over the classes whose iou is not NaN
"""
from __future__ import annotations

import numpy as np

from . import ffi


def jaccard_from_counts(counts: np.ndarray) -> float:
    """counts: [B, 3, NC+1] = inter | true | pred pixel counts per image and class (metrics.py:34-45 restated)."""
    c = np.asarray(counts, np.float64)
    inter, true, pred = c[:, 0], c[:, 1], c[:, 2]
    union = true + pred - inter
    ious = []
    for i in range(c.shape[2]):
        legal = true[:, i] > 0                                  # legal_batches: images that contain class i
        if legal.any():
            ious.append(float(np.mean(inter[legal, i] / union[legal, i])))
    return float(np.mean(ious)) if ious else float('nan')


def jaccard_counts(pred: np.ndarray, gt: np.ndarray, num_classes: int, device: int = 0) -> np.ndarray:
    """uint8 label maps [B, ...] -> int64 counts [B, 3, NC+1] computed on the device."""
    p = np.ascontiguousarray(pred, np.uint8)
    g = np.ascontiguousarray(gt, np.uint8)
    if p.shape != g.shape or p.ndim < 2:
        raise ValueError('pred and gt must be uint8 label maps of the same shape [B, ...]')
    B = p.shape[0]
    n = p.size // B
    dp, dg = ffi.DeviceBuffer.from_numpy(p.reshape(-1), device), ffi.DeviceBuffer.from_numpy(g.reshape(-1), device)
    dc = ffi.DeviceBuffer.from_numpy(np.zeros(B * 3 * (num_classes + 1), np.uint64), device)
    ffi._check(ffi.load_library().dlv3p_op_jaccard_counts(device, dp.ptr, dg.ptr, B, n, num_classes, dc.ptr, None))
    ffi.synchronize(device)
    return dc.download((B, 3, num_classes + 1), np.uint64).astype(np.int64)


def jaccard(pred: np.ndarray, gt: np.ndarray, num_classes: int, device: int = 0) -> float:
    """Jaccard(y_true, y_pred) of the reference for predicted label maps `pred` (= argmax of the model output) and ground truth `gt`."""
    return jaccard_from_counts(jaccard_counts(pred, gt, num_classes, device))
