#!/usr/bin/env python3
"""``python main.py --mode enhance --input_path DIR|FILE ...`` -- same flags as the reference's main.py:25-117."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from retinex_image_enhancement_b200.cli import main  # noqa: E402

if __name__ == "__main__":
    main()
