"""upretinex-b200: B200-native classical hot path of UP-Retinex.

Public surface mirrors the reference's modules:

    retinex_image_enhancement_b200.enhancers.adaptive_params.AdaptiveParameterAdjuster
    retinex_image_enhancement_b200.enhancers.multi_scale.MultiScaleEnhancer
    retinex_image_enhancement_b200.enhancers.content_aware.ContentAwareEnhancer
    retinex_image_enhancement_b200.enhancers.simple_enhance.{enhance_single_image, enhance_batch_images, ...}
    retinex_image_enhancement_b200.losses.loss.{calculate_texture_complexity, dynamic_smooth_weight}
    retinex_image_enhancement_b200.models.model.UP_Retinex

All arithmetic runs in libupretinex_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/upretinex_b200.h).  There is no CPU fallback: importing works anywhere, calling an
op without the built library or without a CUDA device raises.
"""
__version__ = "0.1.0"
