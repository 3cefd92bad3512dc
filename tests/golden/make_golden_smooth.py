#!/usr/bin/env python3
"""Golden vectors for the smoothness term from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_smooth.py

Executes losses/loss.py EdgeAwareSmoothnessLoss.forward (:136-176), AdaptiveExposureLoss (:29-58), ColorLoss (:351-368) and
SpatialConsistencyLoss (:404-427) with torch autograd on seeded inputs and stores the loss values and the gradients w.r.t.
illu_map / img_enhanced in tests/golden/smooth_loss.npz (committed).  Inputs are re-created from the seeds by smooth_cases()
and enh_cases() below, which the tests import.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def smooth_cases():
    """(name, illu [B,Ci,H,W], img_low [B,Cs,H,W], lambda, alpha) -- deterministic."""
    out = []
    rng = np.random.default_rng(71)
    out.append(("a", rng.random((2, 1, 24, 40), dtype=np.float32), rng.random((2, 3, 24, 40), dtype=np.float32), 10.0, 1.0))
    rng = np.random.default_rng(72)
    out.append(("b", rng.random((3, 3, 33, 17), dtype=np.float32), rng.random((3, 3, 33, 17), dtype=np.float32) * np.float32(0.3), 4.0, 0.5))
    rng = np.random.default_rng(73)   # plateaus in illu: exercises sign(0) = 0; single-channel image
    illu = np.floor(rng.random((2, 1, 16, 16), dtype=np.float32) * 4) / np.float32(4)
    out.append(("c", illu.astype(np.float32), rng.random((2, 1, 16, 16), dtype=np.float32), 10.0, 1.0))
    rng = np.random.default_rng(74)   # smallest legal frame
    out.append(("d", rng.random((1, 1, 2, 2), dtype=np.float32), rng.random((1, 3, 2, 2), dtype=np.float32), 10.0, 1.0))
    return out


def enh_cases():
    """(name, enhanced [B,3,H,W], img_low [B,3,H,W]) -- deterministic."""
    out = []
    rng = np.random.default_rng(81)
    out.append(("e1", rng.random((2, 3, 48, 64), dtype=np.float32), rng.random((2, 3, 48, 64), dtype=np.float32) * np.float32(0.3)))
    rng = np.random.default_rng(82)    # ragged: 37 x 53 -> 2 x 3 patches, border pixels outside every patch
    out.append(("e2", rng.random((3, 3, 37, 53), dtype=np.float32) * np.float32(0.8), rng.random((3, 3, 37, 53), dtype=np.float32)))
    rng = np.random.default_rng(83)    # one patch
    out.append(("e3", rng.random((1, 3, 16, 16), dtype=np.float32), rng.random((1, 3, 16, 16), dtype=np.float32)))
    return out


def main():
    import torch
    ref = os.environ.get("UPR_REFERENCE", "/root/reference")
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref)
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    from losses.loss import (AdaptiveExposureLoss, ColorLoss, EdgeAwareSmoothnessLoss,  # noqa: E402  (reference)
                             SpatialConsistencyLoss)
    torch.set_num_threads(1)
    store = {}
    for name, illu, img, lam, alpha in smooth_cases():
        mod = EdgeAwareSmoothnessLoss(lambda_val=lam, alpha=alpha)
        it = torch.from_numpy(illu).clone().requires_grad_(True)
        loss = mod(it, torch.from_numpy(img))
        loss.backward()
        store[f"{name}_loss"] = np.float32(loss.item())
        store[f"{name}_grad"] = it.grad.numpy().astype(np.float32)
        print(name, illu.shape, img.shape, float(loss))
    for name, enh, low in enh_cases():
        lt = torch.from_numpy(low)
        for tag, fn in (("exp", lambda e: AdaptiveExposureLoss()(e, lt)), ("col", lambda e: ColorLoss()(e)),
                        ("spa", lambda e: SpatialConsistencyLoss()(e, lt))):
            et = torch.from_numpy(enh).clone().requires_grad_(True)
            loss = fn(et)
            loss.backward()
            store[f"{name}_{tag}_loss"] = np.float32(loss.item())
            store[f"{name}_{tag}_grad"] = et.grad.numpy().astype(np.float32)
            print(name, tag, float(loss.item()))
    np.savez_compressed(os.path.join(HERE, "smooth_loss.npz"), **store)


if __name__ == "__main__":
    main()
