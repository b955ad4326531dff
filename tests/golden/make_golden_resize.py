"""Golden vectors for the read-time resize of the reference (scipy.misc.imresize -> PIL BILINEAR): outputs of the Pillow
installed in the build container on seeded random uint8 images.  Run once here: `python tests/golden/make_golden_resize.py`
(writes tests/golden/resize_bilinear_golden.npz)."""
import os

import numpy as np
import PIL
from PIL import Image

CASES = [(37, 53, 24, 31), (20, 20, 33, 47), (64, 48, 64, 96), (30, 30, 7, 9), (61, 83, 56, 56), (120, 160, 227, 227)]


def main():
    rng = np.random.default_rng(20261018)
    out = {"pillow_version": np.array(PIL.__version__)}
    for i, (h, w, oh, ow) in enumerate(CASES):
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if i % 2 == 1:  # smooth content as well as noise
            yy, xx = np.mgrid[0:h, 0:w]
            a = np.stack([(yy * 3 + xx) % 256, (xx * 5) % 256, (yy * xx) % 256], axis=-1).astype(np.uint8)
        ref = np.asarray(Image.fromarray(a).resize((ow, oh), resample=Image.BILINEAR))
        out["in_%d" % i] = a
        out["out_%d" % i] = ref
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resize_bilinear_golden.npz"), **out)
    print("wrote %d cases, Pillow %s" % (len(CASES), PIL.__version__))


if __name__ == "__main__":
    main()
