"""Generates tests/golden/digits_bits.npy from the reference's own test fixture
TEST/digitos/digits.png (identical to EXAMPLES/digits.png): a 1600x160 8-bit
gray image whose pixels are all 0 or 255 (1000 handwritten 16x16 digits).

Run in the build container only (needs /root/reference and PIL):
    python tests/golden/make_digits_fixture.py
The fixture stores one bit per pixel (1 = ink): after the reference's
ImageIO.read(...):to_grayscale():invert_colors():matrix() chain
(packages/imaging/libpng/c_src/libpng.cc:161, Image/c_src/floatrgb.h:36) a 255
pixel becomes 0.0f and a 0 pixel becomes 1.0f exactly.
"""
import numpy as np
from PIL import Image

a = np.array(Image.open("/root/reference/TEST/digitos/digits.png"))
assert a.shape == (1600, 160) and set(np.unique(a)) == {0, 255}
bits = np.packbits((a == 0).astype(np.uint8), axis=1)  # [1600, 20]
np.save("tests/golden/digits_bits.npy", bits)
print(bits.shape, bits.nbytes)
