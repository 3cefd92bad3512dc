"""Drop-in for the reference's ``enhancers/simple_enhance.py`` drivers (same function names, argument order,
defaults and output file names ``{stem}_enhanced.png``, ``{stem}_illumination.png``, ``{stem}_comparison.png``).

Reference (file:line relative to the reference tree): load_image :23-62, save_image :65-100, create_comparison
:103-132, enhance_single_image :135-199, enhance_batch_images :202-250.

B200-first differences:
  * everything between ``load_image`` and ``save_image`` stays on the device: one H2D of the input, one D2H of the
    two results (the reference crosses PCIe three times per image, adaptive_params.py:188/:136/:198);
  * ``enhance_batch_images`` groups same-sized images into batches (``batch_size``) so that the classical kernels
    run over many frames per launch, honours ``enable_multi_scale`` / ``enable_content_aware`` (the reference
    silently drops them at :240) and, when launched with one process per GPU (torchrun), shards the file list by
    rank -- frames are independent, there is no collective;
  * ``enhance_single_image`` accepts ``adjuster=`` (the reference's main.py:246 passes it and crashes on the
    missing parameter).
File decoding / PNG encoding are host-side I/O (PIL), outside the hot path.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from ..models.model import UP_Retinex
from ..utils.letterbox import letterbox_tensor
from .adaptive_params import AdaptiveParameterAdjuster
from .content_aware import ContentAwareEnhancer
from .multi_scale import MultiScaleEnhancer

VALID_EXTENSIONS = {".jpg", ".jpeg", ".png", ".bmp", ".tif", ".tiff"}


def load_image(image_path, max_size=None, device=None):
    """-> ([1,3,H,W] f32 tensor in [0,1], (W, H) of the file).  Host tensor like the reference's by default.  With a CUDA
    ``device`` (what the drivers below pass) the decoded uint8 frame is uploaded as is -- 3 instead of 12 bytes per pixel
    over PCIe -- and de-quantised / letterboxed on the device (upr_letterbox_u8_f32; same values, bit for bit)."""
    from PIL import Image
    img = Image.open(image_path).convert("RGB")
    original_size = img.size
    if device is not None and torch.device(device).type == "cuda":
        from .. import native
        from ..utils.letterbox import letterbox_geometry
        u8 = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).unsqueeze(0)
        h, w = u8.shape[1], u8.shape[2]
        if max_size is None:
            (rh, rw), (top, bottom, left, right) = (h, w), (0, 0, 0, 0)
        else:
            (rh, rw), (top, bottom, left, right), _, _ = letterbox_geometry(h, w, max_size, auto=True, scaleup=False)
        dev = torch.device(device)
        with torch.cuda.device(dev):
            t = native.letterbox(u8.pin_memory().to(dev, non_blocking=True), (rh, rw), top, left,
                                 (rh + top + bottom, rw + left + right))
        return t, original_size
    t = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).to(torch.float32) / 255.0
    if max_size is not None:
        t, _, _ = letterbox_tensor(t, new_shape=max_size, auto=True, scaleup=False)
    # max_size None: letterbox to the image's own shape is the identity on a k/255 grid (see utils/letterbox.py)
    return t.unsqueeze(0).contiguous(), original_size


def _to_u8_hwc(tensor):
    """[1,C,H,W] / [C,H,W] f32 -> [H,W,3] uint8 array: (clip(x,0,1)*255).astype(uint8) as in the reference's save_image
    (enhancers/simple_enhance.py:65-80).  A CUDA tensor is clipped, quantised and interleaved on the device, so only 3 bytes
    per pixel cross PCIe (the same truncating cast; uint8 arrays pass through)."""
    if isinstance(tensor, np.ndarray):
        return tensor
    if tensor.dim() == 4:
        tensor = tensor.squeeze(0)
    t = tensor.detach().to(torch.float32)
    if t.is_cuda:
        q = (t.clamp(0, 1) * 255).to(torch.uint8)
        if q.shape[0] == 1:
            q = q.expand(3, -1, -1)
        return q.permute(1, 2, 0).contiguous().cpu().numpy()
    a = (np.clip(t.numpy(), 0, 1) * 255).astype(np.uint8)
    if a.shape[0] == 1:
        return np.stack([a[0]] * 3, axis=2)
    return np.transpose(a, (1, 2, 0))


def save_image(tensor, save_path):
    from PIL import Image
    Image.fromarray(_to_u8_hwc(tensor)).save(save_path)
    print(f"已保存: {save_path}")


def create_comparison(img_low, img_enhanced, save_path):
    from PIL import Image
    Image.fromarray(np.concatenate([_to_u8_hwc(img_low), _to_u8_hwc(img_enhanced)], axis=1)).save(save_path)
    print(f"已保存对比图像: {save_path}")


def _enhance_tensor(model, img_low, device, enable_multi_scale, enable_content_aware, adjuster=None):
    """Dispatch of enhancers/simple_enhance.py:167-175 on a [N,3,H,W] batch."""
    if enable_content_aware:
        return ContentAwareEnhancer().apply_content_aware_enhancement(model, img_low, device)
    if enable_multi_scale:
        return MultiScaleEnhancer().enhance_with_pyramid(model, img_low, device)
    return (adjuster or AdaptiveParameterAdjuster()).apply_adaptive_enhancement(model, img_low, device)


def _write_outputs(img_low, img_enhanced, illu_map, image_path, output_dir, pool=None):
    """The reference's three files per image (enhancers/simple_enhance.py:177-195).  Quantisation happens here (on the
    device for CUDA tensors); with ``pool`` (a ThreadPoolExecutor) the PNG encoding runs on host threads while the GPU
    works on the next batch."""
    os.makedirs(output_dir, exist_ok=True)
    stem = os.path.splitext(os.path.basename(image_path))[0]
    low, enh, illu = _to_u8_hwc(img_low), _to_u8_hwc(img_enhanced), _to_u8_hwc(illu_map)

    def write():
        save_image(enh, os.path.join(output_dir, f"{stem}_enhanced.png"))
        save_image(illu, os.path.join(output_dir, f"{stem}_illumination.png"))
        create_comparison(low, enh, os.path.join(output_dir, f"{stem}_comparison.png"))

    if pool is None:
        write()
        return None
    return pool.submit(write)


def enhance_single_image(model, image_path, output_dir, device, max_size=None, enable_multi_scale=False,
                         enable_content_aware=False, adjuster=None):
    print(f"正在处理: {os.path.basename(image_path)}")
    img_low, _original_size = load_image(image_path, max_size, device=device)
    start = time.time()
    img_enhanced, illu_map = _enhance_tensor(model, img_low, device, enable_multi_scale, enable_content_aware, adjuster)
    if img_enhanced.is_cuda:
        torch.cuda.synchronize(img_enhanced.device)
    print(f"增强耗时: {time.time() - start:.4f}s")
    _write_outputs(img_low, img_enhanced, illu_map, image_path, output_dir)
    print("图像增强完成！")


def list_images(input_dir):
    return sorted(os.path.join(input_dir, f) for f in os.listdir(input_dir)
                  if os.path.splitext(f)[1].lower() in VALID_EXTENSIONS)


def shard_for_rank(items, rank=None, world=None):
    """Contiguous block of ``items`` owned by this rank (frame i -> GPU i*G/N; no collective on the enhance path)."""
    if rank is None or world is None:
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    n = len(items)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    return items[lo:hi]


def enhance_batch_images(input_dir, output_dir, device, max_size=None, enable_multi_scale=False,
                         enable_content_aware=False, batch_size=16, model=None):
    print("正在加载模型...")
    if model is None:
        model = UP_Retinex().to(device).eval()
    files = list_images(input_dir)
    if not files:
        print(f"在目录 '{input_dir}' 中未找到有效图像文件")
        return
    mine = shard_for_rank(files)
    print(f"找到 {len(files)} 个图像文件 (本进程处理 {len(mine)} 个)")
    t0 = time.time()
    pending = []   # consecutive same-shape images form one device batch
    from concurrent.futures import ThreadPoolExecutor
    writers = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1))
    futures = []

    def flush():
        if not pending:
            return
        batch = torch.cat([p[1] for p in pending], dim=0)
        enhanced, illu = _enhance_tensor(model, batch, device, enable_multi_scale, enable_content_aware)
        for i, (path, low) in enumerate(pending):
            futures.append(_write_outputs(low, enhanced[i:i + 1], illu[i:i + 1], path, output_dir, pool=writers))
        pending.clear()

    try:
        for path in mine:
            low, _ = load_image(path, max_size, device=device)
            if pending and (pending[0][1].shape != low.shape or len(pending) >= batch_size):
                flush()
            pending.append((path, low))
        flush()
        for fut in futures:
            fut.result()
    finally:
        writers.shutdown(wait=True)
    total = time.time() - t0
    print("=" * 50)
    print(f"总共处理了 {len(mine)} 张图像")
    print(f"总耗时: {total:.2f}s")
    if mine:
        print(f"平均每张图像耗时: {total / len(mine):.4f}s")
    print("=" * 50)
