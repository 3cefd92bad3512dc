#!/usr/bin/env python3
"""Generate the golden fixtures by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist
on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (committed):
    tests/golden/golden.json      hashes + scalars of the KATs of SURVEY.md section 8c
    tests/golden/small_cases.npz  full arrays for small shapes (bit-exact checks)

Reference functions executed (paths relative to /root/reference):
    enhancers/adaptive_params.py : AdaptiveParameterAdjuster.{apply_clahe_enhancement,
                                   calculate_brightness_features, adjust_parameters}
    enhancers/multi_scale.py     : MultiScaleEnhancer.extract_multi_scale_features (+ :87-94)
    enhancers/content_aware.py   : ContentAwareEnhancer.{compute_saliency_map, compute_attention_map}
    models/model.py              : UP_Retinex.retinex_decompose, recombination formula :442
    losses/loss.py               : calculate_texture_complexity, :710-717
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("UPR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))

from enhancers.adaptive_params import AdaptiveParameterAdjuster  # noqa: E402  (reference)
from enhancers.content_aware import ContentAwareEnhancer  # noqa: E402  (reference)
from enhancers.multi_scale import MultiScaleEnhancer  # noqa: E402  (reference)
from losses.loss import calculate_texture_complexity  # noqa: E402  (reference)
from models.model import UP_Retinex  # noqa: E402  (reference)

from oracle.oracle import kat_input  # noqa: E402  (input generator only)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


CLAHE_KATS = [  # (seed, h, w, kind)
    (1, 400, 600, "uniform"), (2, 400, 600, "dark"), (3, 1080, 1920, "uniform"),
    (4, 1080, 1920, "dark"), (5, 1080, 1920, "ramp"), (6, 2160, 3840, "dark"),
    (7, 403, 601, "uniform"),
]
SMALL = [  # full arrays kept
    (21, 64, 96, "uniform"), (22, 72, 128, "dark"), (23, 67, 93, "uniform"),
    (24, 64, 100, "ramp"), (25, 48, 64, "const"), (26, 128, 256, "dark"),
]


def main():
    torch.manual_seed(0)
    adj, ms, ca = AdaptiveParameterAdjuster(), MultiScaleEnhancer(), ContentAwareEnhancer()
    gold = {"generator": "tests/golden/make_golden.py", "reference": "xh92117/Retinex-image-Enhancement",
            "cv2": __import__("cv2").__version__, "torch": torch.__version__, "numpy": np.__version__}
    small = {}

    gold["clahe"] = []
    for seed, h, w, kind in CLAHE_KATS + SMALL:
        x = kat_input(seed, h, w, kind)
        out = adj.apply_clahe_enhancement(torch.from_numpy(x)).contiguous().numpy()
        q = (np.transpose(x[0], (1, 2, 0)) * 255).astype(np.uint8)
        rec = {"seed": seed, "h": h, "w": w, "kind": kind, "sha_in": sha(x), "sha_u8_hwc": sha(q),
               "sha_out": sha(out), "mean_out": float(out.astype(np.float64).mean())}
        gold["clahe"].append(rec)
        if (seed, h, w, kind) in SMALL:
            small[f"clahe_{seed}_out_u8"] = np.rint(out[0] * 255.0).astype(np.uint8)
            assert np.array_equal(small[f"clahe_{seed}_out_u8"].astype(np.float32) / np.float32(255.0), out[0])

    gold["bright"] = []
    for seed, h, w, kind in [(2, 400, 600, "dark"), (3, 1080, 1920, "uniform"), (21, 64, 96, "uniform"),
                             (5, 1080, 1920, "ramp")]:
        t = torch.from_numpy(kat_input(seed, h, w, kind))
        f = {k: float(v) for k, v in adj.calculate_brightness_features(t).items()}
        gold["bright"].append({"seed": seed, "h": h, "w": w, "kind": kind, "features": f,
                               "params": adj.adjust_parameters(t)})

    gold["multiscale"] = []
    for seed, h, w, kind in [(1, 400, 600, "uniform"), (4, 1080, 1920, "dark"), (7, 403, 601, "uniform"),
                             (23, 67, 93, "uniform"), (24, 64, 100, "ramp")]:
        t = torch.from_numpy(kat_input(seed, h, w, kind))
        feats = ms.extract_multi_scale_features(t)
        means = [float(torch.mean(f).item()) for f in feats]
        factor = 1.0
        for wt, m in zip([0.5, 0.3, 0.2], means):
            factor += wt * m * 0.1
        gold["multiscale"].append({"seed": seed, "h": h, "w": w, "kind": kind, "means": means,
                                   "factor": factor, "shapes": [list(f.shape) for f in feats]})

    gold["content"] = []
    for seed, h, w, kind in [(1, 400, 600, "uniform"), (4, 1080, 1920, "dark"), (23, 67, 93, "uniform"),
                             (22, 72, 128, "dark")]:
        t = torch.from_numpy(kat_input(seed, h, w, kind))
        sal = ca.compute_saliency_map(t).numpy()
        att = ca.compute_attention_map(t).numpy()
        gold["content"].append({"seed": seed, "h": h, "w": w, "kind": kind,
                                "sal_mean": float(sal.astype(np.float64).mean()),
                                "att_mean": float(att.astype(np.float64).mean()),
                                "att_argmax": int(att.argmax()), "sal_argmax": int(sal.argmax())})
        if h * w < 20000:
            small[f"sal_{seed}"] = sal[0, 0]
            small[f"att_{seed}"] = att[0, 0]

    # Retinex arithmetic (model.py:405-413, :442) -- no conv weights involved
    model = UP_Retinex()
    rng = np.random.default_rng(31)
    x = rng.random((2, 3, 24, 40), dtype=np.float32)
    illu = (rng.random((2, 1, 24, 40), dtype=np.float32) * np.float32(0.9) + np.float32(0.05))
    illu[0, 0, 0, :4] = 0.0  # exercise the epsilon
    e = rng.random((2, 3, 24, 40), dtype=np.float32)
    refl = model.retinex_decompose(torch.from_numpy(x), torch.from_numpy(illu))
    et = torch.from_numpy(e)
    enh = refl * et + (1 - refl) * (et ** 2)
    small["retinex_x"], small["retinex_illu"], small["retinex_e"] = x, illu, e
    small["retinex_refl"], small["retinex_enh"] = refl.numpy(), enh.numpy()

    gold["texture"] = []
    for seed, kind, shape in [(11, "uniform", (8, 3, 256, 256)), (12, "dark", (8, 3, 256, 256)),
                              (13, "uniform", (3, 3, 37, 53)), (14, "uniform", (2, 1, 40, 40))]:
        rng = np.random.default_rng(seed)
        a = rng.random(shape, dtype=np.float32)
        if kind == "dark":
            a = a * np.float32(0.3)
        t = torch.from_numpy(a)
        tv = calculate_texture_complexity(t, "tv")
        ed = calculate_texture_complexity(t, "edge_density")
        w_tv = torch.clamp(1.0 * (1.0 - torch.mean(tv) * 0.8), 0.1, 5.0)
        w_ed = torch.clamp(1.0 * (1.0 - torch.mean(ed) * 0.8), 0.1, 5.0)
        gold["texture"].append({"seed": seed, "kind": kind, "shape": list(shape), "sha_in": sha(a),
                                "tv": [float(v) for v in tv], "edge_density": [float(v) for v in ed],
                                "w_tv": float(w_tv), "w_edge": float(w_ed)})

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **small)
    print("wrote golden.json and small_cases.npz:", {k: (len(v) if isinstance(v, list) else v) for k, v in gold.items()})


if __name__ == "__main__":
    main()
