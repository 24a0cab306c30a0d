"""Generates tests/golden/shadows_reference_128.npy from the reference's published render
docs/src/assets/shadows-sppm-1024x1024_mio.png ("Output from /scenes/shadows.jl", README.md:9-11): the 1024x1024 8-bit
RGB image box-filtered to 128x128 (float32 in [0, 1]).  Run in the build container (needs /root/reference and PIL):
    python scripts/make_golden_shadows.py
The fixture travels to the GPU box; /root/reference does not."""
import os, sys
import numpy as np
from PIL import Image
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = "/root/reference/docs/src/assets/shadows-sppm-1024x1024_mio.png"
img = np.asarray(Image.open(src).convert("RGB"), dtype=np.float32) / 255.0
assert img.shape == (1024, 1024, 3), img.shape
small = img.reshape(128, 8, 128, 8, 3).mean(axis=(1, 3)).astype(np.float32)
out = os.path.join(ROOT, "tests", "golden", "shadows_reference_128.npy")
np.save(out, (small * 255.0 + 0.5).astype(np.uint8))
print("wrote", out, small.shape, "mean", float(small.mean()))
