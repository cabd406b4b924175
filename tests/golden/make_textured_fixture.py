#!/usr/bin/env python
"""Writes tests/golden/textured/{quads.obj,quads.mtl,pattern.png}: a small synthetic textured mesh for the load_obj /
texture-bake tests (own content, generated -- nothing is taken from the reference's data files).

Two quads (fan-triangulated by the loader) and one triangle: material `painted` has a map_Kd image and UVs that include
values above 1 (wrapped by the loader, load_obj.py:66) and exactly 1 (the reference's one-past-the-edge taps);
material `flat` has only a Kd colour; the first face has no material and no UVs (0.5 grey / vt index 0 quirk)."""
import os

import numpy as np
from PIL import Image

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "textured")


def main():
    os.makedirs(HERE, exist_ok=True)
    h, w = 64, 48
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(x * 5 + y) % 256, (y * 3 + 2 * x) % 256, ((x // 6 + y // 8) % 2) * 200 + 20], axis=-1).astype(np.uint8)
    Image.fromarray(img, "RGB").save(os.path.join(HERE, "pattern.png"))
    with open(os.path.join(HERE, "quads.mtl"), "w") as f:
        f.write("newmtl painted\nKd 0.2 0.4 0.6\nmap_Kd pattern.png\n\nnewmtl flat\nKd 0.9 0.1 0.3\n")
    with open(os.path.join(HERE, "quads.obj"), "w") as f:
        f.write("mtllib quads.mtl\n")
        for v in [(-1, -1, 0), (1, -1, 0.2), (1, 1, 0), (-1, 1, -0.2), (0, 0, 1.5), (2, 0.5, 1.0), (2, -0.5, 0.5)]:
            f.write("v %g %g %g\n" % v)
        for vt in [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (0.31, 0.77), (1.6, 0.25), (0.5, 2.4)]:
            f.write("vt %g %g\n" % vt)
        f.write("f 5 6 7\n")
        f.write("usemtl painted\nf 1/1 2/2 3/3 4/4\nf 1/5 2/6 5/7\n")
        f.write("usemtl flat\nf 2/2 6/6 7/7 3/3\n")


if __name__ == "__main__":
    main()
