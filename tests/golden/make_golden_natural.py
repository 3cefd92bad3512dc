#!/usr/bin/env python3
"""Natural-image golden fixtures: crops of the reference's own sample photographs run through the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_natural.py

Every parity input of round 1 was synthetic (noise / ramps / constants).  The reference ships five JPEGs
(/root/reference/data/input/); this script decodes them the way the reference's own driver does
(enhancers/simple_enhance.py:36-40: PIL -> RGB -> ToTensor, i.e. u8 / 255 in fp32), cuts crops small enough to commit,
and records what the reference computes on them:

    tests/golden/natural.npz    the u8 crops (HWC) + sub-sampled saliency / attention maps (every 4th pixel of every 4th row)
    tests/golden/natural.json   CLAHE hashes (a1), brightness features (a3), multi-scale means / factor (a4/a5),
                                saliency / attention means, extrema positions (a6/a7), texture statistics (a9)

Reference functions executed (paths relative to /root/reference): enhancers/adaptive_params.py
AdaptiveParameterAdjuster.{apply_clahe_enhancement, calculate_brightness_features, adjust_parameters};
enhancers/multi_scale.py MultiScaleEnhancer.extract_multi_scale_features (+ :87-94); enhancers/content_aware.py
ContentAwareEnhancer.{compute_saliency_map, compute_attention_map}; losses/loss.py calculate_texture_complexity.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch
from PIL import Image

REF = os.environ.get("UPR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))

from enhancers.adaptive_params import AdaptiveParameterAdjuster  # noqa: E402  (reference)
from enhancers.content_aware import ContentAwareEnhancer  # noqa: E402  (reference)
from enhancers.multi_scale import MultiScaleEnhancer  # noqa: E402  (reference)
from losses.loss import calculate_texture_complexity  # noqa: E402  (reference)

# (name, file, top, left, height, width): one 8x8-tile-friendly crop (vector kernels), one ragged crop (generic kernels, the
# OpenCV padding quirk, odd pyramid sizes), one very dark crop
CROPS = [
    ("road_stripe", "094216845-003241-003241.jpg", 96, 512, 384, 512),
    ("asphalt_ragged", "102904222-004389-004389.jpg", 301, 222, 250, 333),
    ("dark_edge", "102959263-004697-004697.jpg", 640, 0, 256, 320),
]
SUB = 4


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    adj, ms, ca = AdaptiveParameterAdjuster(), MultiScaleEnhancer(), ContentAwareEnhancer()
    gold = {"generator": "tests/golden/make_golden_natural.py", "reference": "xh92117/Retinex-image-Enhancement",
            "cv2": __import__("cv2").__version__, "torch": torch.__version__, "numpy": np.__version__, "sub": SUB, "cases": []}
    arrays = {}
    for name, fn, top, left, h, w in CROPS:
        img = np.asarray(Image.open(os.path.join(REF, "data", "input", fn)).convert("RGB"))
        crop = np.ascontiguousarray(img[top:top + h, left:left + w])
        assert crop.shape == (h, w, 3)
        arrays[f"{name}_u8"] = crop
        # ToTensor(): u8 HWC -> f32 CHW / 255 (torchvision divides by 255 in fp32)
        x = torch.from_numpy(crop).permute(2, 0, 1).contiguous().to(torch.float32).div(255).unsqueeze(0)
        rec = {"name": name, "file": fn, "top": top, "left": left, "h": h, "w": w, "sha_u8": sha(crop), "sha_in": sha(x.numpy())}
        out = adj.apply_clahe_enhancement(x).contiguous().numpy()
        rec["clahe"] = {"sha_out": sha(out), "mean_out": float(out.astype(np.float64).mean())}
        rec["bright"] = {"features": {k: float(v) for k, v in adj.calculate_brightness_features(x).items()},
                         "params": adj.adjust_parameters(x)}
        feats = ms.extract_multi_scale_features(x)
        means = [float(torch.mean(f).item()) for f in feats]
        factor = 1.0
        for wt, m in zip([0.5, 0.3, 0.2], means):
            factor += wt * m * 0.1
        rec["multiscale"] = {"means": means, "factor": factor, "shapes": [list(f.shape) for f in feats]}
        sal = ca.compute_saliency_map(x).numpy()
        att = ca.compute_attention_map(x).numpy()
        rec["content"] = {"sal_mean": float(sal.astype(np.float64).mean()), "att_mean": float(att.astype(np.float64).mean()),
                          "sal_argmax": int(sal.argmax()), "att_argmax": int(att.argmax()),
                          "sal_max": float(sal.max()), "att_max": float(att.max()), "sal_min": float(sal.min()), "att_min": float(att.min())}
        arrays[f"{name}_sal_sub"] = np.ascontiguousarray(sal[0, 0, ::SUB, ::SUB])
        arrays[f"{name}_att_sub"] = np.ascontiguousarray(att[0, 0, ::SUB, ::SUB])
        rec["texture"] = {"tv": float(calculate_texture_complexity(x, "tv")[0]),
                          "edge_density": float(calculate_texture_complexity(x, "edge_density")[0])}
        gold["cases"].append(rec)
    with open(os.path.join(HERE, "natural.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "natural.npz"), **arrays)
    print("wrote natural.json / natural.npz:", [(c["name"], c["h"], c["w"]) for c in gold["cases"]],
          os.path.getsize(os.path.join(HERE, "natural.npz")), "bytes")


if __name__ == "__main__":
    main()
