"""Drop-in for the texture-statistics part of the reference's ``losses/loss.py``.

Reference behaviour mirrored (file:line relative to the reference tree):
  * calculate_texture_complexity   losses/loss.py:523-583   ('tv' and 'edge_density', ValueError otherwise)
  * dynamic smoothness weight      losses/loss.py:704-720   clamp(w0 * (1 - 0.8 * mean_B(c)), 0.1, 5.0)

The seven differentiable loss terms (loss.py:12-520) need autograd and stay stock PyTorch: they are out of
scope of the hot path.  ``DynamicSmoothWeight`` is the piece ``TotalLoss.forward`` would call at :707-717;
under data parallelism (one process per GPU, torch.distributed initialised) the batch mean of :710 becomes
ONE all-reduce(SUM) of the two floats [sum of complexity, image count], so every rank derives the same
weight as a single process would on the concatenated batch.
"""
from __future__ import annotations

import torch

from .. import native


def calculate_texture_complexity(img: torch.Tensor, method: str = "tv") -> torch.Tensor:
    """[B,C,H,W] f32 CUDA -> [B] f32 CUDA.  Same error behaviour as the reference for unknown methods."""
    if method not in ("tv", "edge_density"):
        raise ValueError(f"不支持的纹理复杂度计算方法: {method}")
    return native.texture_complexity(img, method)


def batch_texture_stats(img: torch.Tensor, method: str = "tv"):
    """(per-image complexity [B], [sum, B] f32 pair ready for the all-reduce)."""
    if method not in ("tv", "edge_density"):
        raise ValueError(f"不支持的纹理复杂度计算方法: {method}")
    return native.texture_complexity(img, method, want_batch_stats=True)


def all_reduce_batch_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce(SUM) of the [sum, count] pair when a process group exists; identity otherwise."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def weight_from_stats(stats: torch.Tensor, weight_smooth: float = 1.0) -> torch.Tensor:
    """0-dim tensor clamp(w0 * (1 - 0.8 * stats[0]/stats[1]), 0.1, 5.0); CUDA kernel for CUDA stats, and the
    same three fp32 operations in torch for host-side (gloo) tensors."""
    if stats.is_cuda:
        return native.dynamic_smooth_weight(stats, weight_smooth)
    avg = stats[0] / stats[1]
    return torch.clamp(torch.tensor(weight_smooth, dtype=torch.float32) * (1.0 - avg * 0.8), 0.1, 5.0)


class DynamicSmoothWeight:
    """``TotalLoss``'s dynamic smoothness weight (losses/loss.py:607-656 constructor arguments
    ``weight_smooth``, ``use_dynamic_smooth_weight``, ``texture_method``)."""

    def __init__(self, weight_smooth: float = 1.0, use_dynamic_smooth_weight: bool = True, texture_method: str = "tv",
                 group=None):
        self.weight_smooth = weight_smooth
        self.use_dynamic_smooth_weight = use_dynamic_smooth_weight
        self.texture_method = texture_method
        self.group = group

    def __call__(self, img_low: torch.Tensor) -> torch.Tensor:
        if not self.use_dynamic_smooth_weight:
            return torch.tensor(self.weight_smooth, dtype=torch.float32, device=img_low.device)
        _per_image, stats = batch_texture_stats(img_low, self.texture_method)
        all_reduce_batch_stats(stats, self.group)
        return weight_from_stats(stats, self.weight_smooth)
