#!/usr/bin/env python3
"""``python simple_enhance.py --input X [--multi_scale] [--content_aware]`` -- flags of the reference's simple_enhance.py:17-42."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from retinex_image_enhancement_b200.cli import simple_enhance_main  # noqa: E402

if __name__ == "__main__":
    simple_enhance_main()
