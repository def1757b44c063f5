"""Drop-in for testing/metrics.py: `Recall`, `Precision`, `F1score` (nn.Module, same call signature) and `psnr`.

The reference walks the non-zero positions of one mask in Python and multiplies a 3x3 window of the other
mask by a kernel that is zero except for its centre (metrics.py:30, :109), i.e. a peak looks at ONE position of
the other mask: its own, except one step further at index 0 of either axis, where window and kernel are sliced
from opposite ends (metrics.py:44-47, :72-75; reproduced).  Here both sums come from one pass on the GPU
(`mfpa_mask_metrics`).  `psnr` reproduces `torchmetrics.PeakSignalNoiseRatio(average="micro")` as the module
configures it (metrics.py:7; torchmetrics is a third-party dependency absent from the reference tree and this
image - restated from its published definition, parity unpinned): data range = max(target, 0) - min(target, 0)
of the call, 10 log10(range^2 / mean squared error).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _ctx():
    from musicfpaugment_b200 import runtime

    return runtime.get_context()


def _mask_sums(predicted: torch.Tensor, gt: torch.Tensor):
    ctx = _ctx()
    dev = torch.device("cuda", ctx.device)
    if predicted.shape != gt.shape or predicted.dim() != 3:
        raise RuntimeError(f"masks must be [batch, n, m] of equal shape: {tuple(predicted.shape)} vs {tuple(gt.shape)}")
    s = ctx.mask_metrics(predicted.to(dev, torch.float32), gt.to(dev, torch.float32)).cpu().tolist()
    return s  # n_gt, pred@gt, n_pred, gt@pred


class Recall(nn.Module):
    """metrics.py:10-85: sum of `predicted` at the positions of the ground-truth peaks / number of ground-truth peaks."""

    def forward(self, predicted: torch.Tensor, gt: torch.Tensor, device: str = "cpu") -> float:
        n_gt, hit, _, _ = _mask_sums(predicted, gt)
        return 0.0 if n_gt == 0 else float(hit / n_gt)


class Precision(nn.Module):
    """metrics.py:88-162: sum of `gt` at the positions of the predicted peaks / number of predicted peaks."""

    def forward(self, predicted: torch.Tensor, gt: torch.Tensor, device: str = "cpu") -> float:
        _, _, n_pred, hit = _mask_sums(predicted, gt)
        return 0.0 if n_pred == 0 else float(hit / n_pred)


class F1score(nn.Module):
    """metrics.py:165-192."""

    def __init__(self) -> None:
        super().__init__()
        self.prec = Precision()
        self.rec = Recall()

    def forward(self, predicted: torch.Tensor, gt: torch.Tensor, device: str = "cpu") -> float:
        n_gt, hit_r, n_pred, hit_p = _mask_sums(predicted, gt)
        p = 0.0 if n_pred == 0 else hit_p / n_pred
        r = 0.0 if n_gt == 0 else hit_r / n_gt
        if math.isclose(p + r, 0.0):
            return 0.0
        return float(2.0 * (p * r) / (p + r))


class _PSNR:
    """Callable like the module-level `psnr = PeakSignalNoiseRatio(average="micro")` (metrics.py:7): returns a
    0-d tensor (callers take `.item()`, audfprint_exps.py:131-132)."""

    def __call__(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        ctx = _ctx()
        dev = torch.device("cuda", ctx.device)
        if preds.shape != target.shape:
            raise RuntimeError(f"Predictions and targets are expected to have the same shape, got {tuple(preds.shape)} and {tuple(target.shape)}")
        sse, tmin, tmax = ctx.psnr_stats(preds.to(dev, torch.float64), target.to(dev, torch.float64)).cpu().tolist()
        data_range = max(tmax, 0.0) - min(tmin, 0.0)   # torchmetrics tracks min / max of the target starting from 0
        mse = sse / target.numel()
        return torch.tensor(10.0 * math.log10(data_range ** 2 / mse) if mse > 0 else float("inf"))


psnr = _PSNR()
