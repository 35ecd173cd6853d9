"""Device-side evaluation metrics for the Siamese change-detection model.

Mirrors ``calculate_metrics`` (evaluate.py:34-64) and the per-sample accumulation of ``evaluate_model``
(evaluate.py:129-210): the reference applies ``torch.sigmoid``, moves every prediction map to the CPU and thresholds it
there, one sample at a time.  Here one kernel (``gap_seg_confusion``) turns a batch of logits into per-sample
[TP, FP, FN, TN] counts on the GPU; only 4 integers per sample cross PCIe, and the ratios are formed from them with the
reference's formulas (fp32 arithmetic, ``smooth`` in the same places).
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import ops

METRIC_KEYS = ("accuracy", "precision", "recall", "f1", "iou")


def confusion_counts(logits: torch.Tensor, labels: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Per-sample int64 [n, 4] = [TP, FP, FN, TN] of ``sigmoid(logits) > 0.5`` against {0,1} ``labels`` (int64 [n,H,W]
    as the dataset yields them, or fp32).  ``out`` (zeroed by the caller) is accumulated into when given."""
    n = logits.shape[0]
    if out is None:
        out = torch.zeros(n, 4, device=logits.device, dtype=torch.int64)
    ops.seg_confusion(logits.contiguous(), labels.contiguous(), out)
    return out


def metrics_from_counts(tp, fp, fn, tn, smooth: float = 1e-6) -> Dict[str, float]:
    """The ratios of evaluate.py:47-64 from the four counts (fp32 arithmetic like the reference's tensors)."""
    f = torch.float32
    tp, fp, fn, tn = (torch.as_tensor(float(v), dtype=f) for v in (tp, fp, fn, tn))
    precision = (tp + smooth) / (tp + fp + smooth)
    recall = (tp + smooth) / (tp + fn + smooth)
    f1 = (2 * precision * recall + smooth) / (precision + recall + smooth)
    union = (tp + fp) + (tp + fn) - tp          # preds.sum() + targets.sum() - intersection
    iou = (tp + smooth) / (union + smooth)
    accuracy = (tp + tn + smooth) / (tp + tn + fp + fn + smooth)
    return {"accuracy": accuracy.item(), "precision": precision.item(), "recall": recall.item(), "f1": f1.item(),
            "iou": iou.item()}


def calculate_metrics(logits: torch.Tensor, targets: torch.Tensor, smooth: float = 1e-6) -> Dict[str, float]:
    """evaluate.py:34 for ONE sample or a whole flattened batch, from logits (the reference passes sigmoid(logits))."""
    c = confusion_counts(logits.reshape(1, -1), targets.reshape(1, -1)).cpu()[0]
    return metrics_from_counts(*c.tolist(), smooth=smooth)


def batch_metrics(logits: torch.Tensor, labels: torch.Tensor, smooth: float = 1e-6) -> List[Dict[str, float]]:
    """Per-sample metric dicts of a batch (the inner loop of evaluate.py:150-185) with one kernel and one small D2H."""
    c = confusion_counts(logits, labels).cpu()
    return [metrics_from_counts(*row.tolist(), smooth=smooth) for row in c]
