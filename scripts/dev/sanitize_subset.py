#!/usr/bin/env python3
"""Small-shape invocation of every kernel family through the C ABI, each checked against the oracle -- the workload that
scripts/dev/sanitize.sh runs under compute-sanitizer (memcheck / racecheck / synccheck / initcheck).  Shapes are small on
purpose: the sanitizer slows kernels down 10-100x."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from retinex_image_enhancement_b200 import extensions as E  # noqa: E402
from retinex_image_enhancement_b200 import native  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main():
    rng = np.random.default_rng(1)
    # a1: vector path (strips + ticket hand-over, column passes), generic path (padding quirk), packed u8, fused Retinex
    for (n, h, w, tiles) in ((2, 64, 96, (8, 8)), (1, 136, 256, (8, 8)), (1, 67, 93, (8, 8)), (1, 32, 1200, (1, 1)), (2, 48, 152, (2, 2))):
        x = np.concatenate([O.kat_input(10 + i, h, w, ("uniform", "dark", "const")[i % 3]) for i in range(n)])
        out = native.clahe_lab(dev(x), 2.0, tiles).cpu().numpy()
        for i in range(n):
            assert np.array_equal(out[i:i + 1], O.clahe_lab(x[i:i + 1], 2.0, tiles)), ("clahe", n, h, w)
        x8 = dev((x * 255).astype(np.uint8).transpose(0, 2, 3, 1))
        o8 = native.clahe_lab_u8(x8, 2.0, tiles)
        assert torch.equal(native.clahe_lab_f32_u8(x8.permute(0, 3, 1, 2).float().div(255).contiguous(), 2.0, tiles), o8)
        e = rng.random((n, 3, h, w), dtype=np.float32)
        illu = (rng.random((n, 1, h, w), dtype=np.float32) * 0.9 + 0.05).astype(np.float32)
        got = native.retinex_clahe(dev(x), dev(illu), dev(e), 2.0, tiles)
        _, e_ref = O.retinex_recombine(x, illu, e)
        for i in range(n):
            assert np.array_equal(got[i:i + 1].cpu().numpy(), O.clahe_lab(e_ref[i:i + 1], 2.0, tiles)), ("fused", n, h, w)
        native.retinex_clahe_u8(dev(x), dev(illu), dev(e), 2.0, tiles)
    # a3, a4/a5 (stream + generic), a6/a7, chained epilogue
    for (n, h, w) in ((2, 64, 120), (1, 72, 244), (1, 33, 51)):
        x = np.concatenate([O.kat_input(30 + i, h, w, ("uniform", "dark")[i % 2]) for i in range(n)])
        xd = dev(x)
        assert np.array_equal(native.brightness_hist(xd).cpu().numpy()[0].astype(np.uint32), O.brightness_hist(x[:1]))
        for force in (False, True):
            m, g = native.multiscale_stats(xd, force_generic=force)
            np.testing.assert_allclose(m.cpu().numpy()[0], O.multiscale_means(x[:1])[0], rtol=2e-6)
        native.multiscale_features(xd)
        np.testing.assert_allclose(native.saliency(xd).cpu().numpy()[:1], O.saliency(x[:1]), rtol=0, atol=1e-6)
        np.testing.assert_allclose(native.attention(xd).cpu().numpy()[:1], O.attention(x[:1]), rtol=0, atol=2e-6)
        enh = dev(rng.random((n, 3, h, w), dtype=np.float32))
        native.content_aware_apply(xd, enh, want_attention=True)
        native.content_multiscale_apply(xd, enh)
        native.quantize_u8(enh)
        native.quantize_u8(enh[:, :1].contiguous())
    # a8, a9/a10 (tv, edge density: ticket hand-over + batch ticket), N3 losses, N2 letterbox, N4 extension ops (TMA + mbarrier)
    a = rng.random((5, 3, 64, 96), dtype=np.float32)
    ad = dev(a)
    np.testing.assert_allclose(native.texture_complexity(ad, "tv").cpu().numpy(), O.texture_tv(a), rtol=2e-6)
    per, stats = native.texture_complexity(ad, "edge_density", want_batch_stats=True)
    native.dynamic_smooth_weight(stats)
    native.texture_weight_peer(ad, "tv", 1.0, None, 0, 1, 1)
    illu = dev(rng.random((5, 1, 64, 96), dtype=np.float32))
    native.retinex_recombine(ad, illu, ad)
    native.edge_smooth_loss(illu, ad)
    losses, saved = native.enhanced_image_losses(ad, ad * 0.5)
    native.enhanced_image_losses_grad(ad, (ad * 0.5).contiguous(), saved, torch.ones(3, device="cuda"))
    u8 = dev(rng.integers(0, 255, (1, 96, 128, 3), dtype=np.uint8))
    native.letterbox(u8, (48, 64), 8, 0, (64, 64))
    xe = dev(rng.random((1, 3, 72, 160), dtype=np.float32) * 0.9 + 0.05)
    E.gaussian_blur(xe, 15)
    E.multi_scale_retinex(xe, (7, 15))
    E.pyr_down(xe)
    E.gamma_correct(xe, 0.45)
    torch.cuda.synchronize()
    print("sanitize subset ok")


if __name__ == "__main__":
    main()
