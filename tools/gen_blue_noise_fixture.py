#!/usr/bin/env python
"""Writes tests/golden/blue_noise_ldr_rgba_64.npz: the 64 blue-noise slices the reference binds to directLight.rgen /
reflection.rgen (data/BlueNoise/64_64/LDR_RGBA_{0..63}.png, reference src/VulkanLifecycle.cpp:113-119), decoded by the
product's own PNG decoder (vkx_image_decode, byte-identical to the reference's stb_image on tests/golden/stb_pin.json).
The fixture travels to the GPU box, where /root/reference does not exist."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vulkanexp_b200._lib import image_decode  # noqa: E402

REF = os.environ.get("VKX_REFERENCE", "/root/reference")

if __name__ == "__main__":
    slices = []
    for i in range(64):
        img = image_decode(os.path.join(REF, "data", "BlueNoise", "64_64", "LDR_RGBA_%d.png" % i))
        assert img.shape == (64, 64, 4), img.shape
        slices.append(img)
    arr = np.stack(slices)
    path = os.path.join(ROOT, "tests", "golden", "blue_noise_ldr_rgba_64.npz")
    np.savez_compressed(path, rgba8=arr)
    print("wrote", path, os.path.getsize(path), "bytes; mean", arr.mean(axis=(0, 1, 2)))
