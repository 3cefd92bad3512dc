"""Model boundary of the hot path: ``UP_Retinex.forward(x) -> (enhanced, reflectance, illu)``.

The CNN of the reference (IENet U-Net, EnhancedFAM, ASPP ..., /root/reference/models/model.py:11-403) is dense
convolution work that cuDNN already covers; it is OUT OF SCOPE of this package.  What
is on the path are the two pointwise Retinex lines around it:

    reflectance = x / (illu + 1e-6)                    models/model.py:405-413  (retinex_decompose)
    enhanced    = R * e + (1 - R) * e**2               models/model.py:442

Both run in one fused sm_100a kernel (upr_retinex_recombine_f32) when the model is in inference mode
(``torch.no_grad()`` / ``eval()``); with autograd enabled the same two lines are evaluated by stock torch ops so
that training still back-propagates through them.

``UP_Retinex`` below is a drop-in for the *interface* of the reference class (constructor flags ``use_preact``,
``use_aspp``; ``forward``; ``retinex_decompose``), with a deliberately small stock-PyTorch stand-in for the
out-of-scope convolution stacks -- the enhance entry points run a randomly initialised network anyway
(enhancers/simple_enhance.py:214-216 loads no checkpoint).  To run the reference's own CNN with the fused
Retinex arithmetic, pass an instance of the reference class to ``accelerate_reference_model``; checkpoints written
by the reference trainer only load into that class.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import native

EPSILON = 1e-6   # models/model.py:411


def retinex_decompose(x: torch.Tensor, illu: torch.Tensor) -> torch.Tensor:
    """R = x / (illu + 1e-6).  Inference: the sm_100a kernel (CUDA tensors required -- there is no CPU path).
    Only when autograd must flow through the line (training) is it left to stock torch ops."""
    if torch.is_grad_enabled() and (x.requires_grad or illu.requires_grad):
        return x / (illu + EPSILON)
    return native.retinex_decompose(x, illu, EPSILON)


def retinex_recombine(x: torch.Tensor, illu: torch.Tensor, enhancement_map: torch.Tensor):
    """(reflectance, enhanced) of models/model.py:412 and :442."""
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or illu.requires_grad or enhancement_map.requires_grad)
    if needs_grad:   # training: autograd through the two lines, stock torch ops
        reflectance = x / (illu + EPSILON)
        return reflectance, reflectance * enhancement_map + (1 - reflectance) * (enhancement_map ** 2)
    return native.retinex_recombine(x, illu, enhancement_map, want_reflectance=True, eps=EPSILON)


class _ConvStack(nn.Module):
    def __init__(self, cin, cmid, cout):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cmid, 3, padding=1)
        self.c2 = nn.Conv2d(cmid, cmid, 3, padding=1)
        self.c3 = nn.Conv2d(cmid, cout, 1)

    def forward(self, x):
        return self.c3(F.relu(self.c2(F.relu(self.c1(x)))))


class MultiScaleUP_Retinex(nn.Module):
    """Interface-compatible stand-in (see module docstring).  ``use_preact`` / ``use_aspp`` are accepted for CLI
    compatibility (main.py:227-229) and do not change this stand-in."""

    def __init__(self, use_preact: bool = True, use_aspp: bool = True, width: int = 16):
        super().__init__()
        self.use_preact, self.use_aspp = use_preact, use_aspp
        self.ie_net = _ConvStack(3, width, 1)            # illumination estimate, 1 channel
        self.scale1 = _ConvStack(3, width, width)        # enhancement-map branches at scales 1, 1/2, 1/4
        self.scale2 = _ConvStack(3, width, width)
        self.scale3 = _ConvStack(3, width, width)
        self.output_layer = nn.Conv2d(3 * width, 3, 1)

    def retinex_decompose(self, x, illu):
        return retinex_decompose(x, illu)

    def enhancement_map(self, x):
        size = x.shape[2:]
        f1 = self.scale1(x)
        f2 = self.scale2(F.interpolate(x, scale_factor=0.5, mode="bilinear", align_corners=False))
        f3 = self.scale3(F.interpolate(x, scale_factor=0.25, mode="bilinear", align_corners=False))
        f2 = F.interpolate(f2, size=size, mode="bilinear", align_corners=False)
        f3 = F.interpolate(f3, size=size, mode="bilinear", align_corners=False)
        return torch.sigmoid(self.output_layer(torch.cat([f1, f2, f3], dim=1)))

    def forward_maps(self, x):
        """The two CNN outputs the Retinex arithmetic consumes: (illumination [B,1,H,W], enhancement map [B,3,H,W]).
        Inference drivers that only need the CLAHE'd result hand them to the fused ``native.retinex_clahe`` instead of
        materialising reflectance and enhanced (``AdaptiveParameterAdjuster.apply_adaptive_enhancement``)."""
        return torch.sigmoid(self.ie_net(x)), self.enhancement_map(x)

    def forward(self, x):
        illu, e = self.forward_maps(x)
        reflectance, enhanced = retinex_recombine(x.contiguous(), illu.contiguous(), e.contiguous())
        return enhanced, reflectance, illu


UP_Retinex = MultiScaleUP_Retinex   # same alias as models/model.py:459


def accelerate_reference_model(model: nn.Module) -> nn.Module:
    """Route ``retinex_decompose`` of an instance of the REFERENCE class through the fused kernel at inference.
    (The recombination line :442 sits inside the reference's ``multi_scale_enhance`` after its convolutions and
    cannot be swapped without re-stating that method; use ``retinex_recombine`` when building on this package.)"""
    model.retinex_decompose = retinex_decompose   # instance attribute shadows the bound method
    return model
