#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference's own test data.

Run HERE (the container that has /root/reference); the GPU box has no reference
tree, so tests read only the small files this script writes next to itself:

  teapot.npz            vertices [1292,3] f32 (unit-cube normalised as load_obj.py:188-192 does) and
                        faces [2464,3] i32 of tests/data/teapot.obj      (test_load_obj.py:34-37)
  golden_images.npz     silhouette   : bit-packed `teapot_blender.png.min(-1) != 255`
                                       (test_rasterize_silhouettes.py:29-33, test_rasterize.py:68-72,
                                        test_rasterize_depth.py:31-35)
                        depth_u8     : tests/data/test_depth.png          (test_rasterize_depth.py:39-58)
                        rasterize1_u8, rasterize2_u8 : the un-asserted snapshots test_rasterize.py:15-50 writes
  kat.json              the two known-answer gradient cases of test_rasterize_silhouettes.py:37-99 /
                        test_rasterize.py:76-149 (numbers are test facts quoted from those files)
"""
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NR_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import nr_oracle  # noqa: E402  (only its numpy OBJ reader is used here)


def main():
    data = os.path.join(REF, "tests", "data")
    vertices, faces = nr_oracle.load_obj(os.path.join(data, "teapot.obj"))
    assert vertices.shape == (1292, 3) and faces.shape == (2464, 3)
    np.savez_compressed(os.path.join(HERE, "teapot.npz"), vertices=vertices, faces=faces)

    blender = np.array(Image.open(os.path.join(data, "teapot_blender.png"))).astype(np.float32)
    sil = (blender.min(-1) != 255)
    depth = np.array(Image.open(os.path.join(data, "test_depth.png")))
    r1 = np.array(Image.open(os.path.join(data, "test_rasterize1.png")))
    r2 = np.array(Image.open(os.path.join(data, "test_rasterize2.png")))
    np.savez_compressed(os.path.join(HERE, "golden_images.npz"), silhouette=np.packbits(sil), depth_u8=depth,
                        rasterize1_u8=r1, rasterize2_u8=r2)

    kat = {
        "source": "tests/test_rasterize_silhouettes.py:37-99 and tests/test_rasterize.py:76-149",
        "image_size": 64, "anti_aliasing": False, "perspective": False, "batch_size": 4, "target_num": 2,
        "rtol": 1e-2,
        "cases": [
            {"name": "out_of_face", "vertices": [[0.8, 0.8, 1.0], [0.0, -0.5, 1.0], [0.2, -0.4, 1.0]],
             "faces": [[0, 1, 2]], "pxi": 35, "pyi": 25, "loss": "sum(abs(image[:, pyi, pxi] - 1))",
             "grad_ref": [[1.6725862, -0.26021874, 0.0], [1.41986704, -1.64284933, 0.0], [0.0, 0.0, 0.0]]},
            {"name": "on_face", "vertices": [[0.8, 0.8, 1.0], [-0.5, -0.8, 1.0], [0.8, -0.8, 1.0]],
             "faces": [[0, 1, 2]], "pxi": 50, "pyi": 40, "loss": "sum(abs(image[:, pyi, pxi]))",
             "grad_ref": [[0.98646867, 1.04628897, 0.0], [-1.03415668, -0.10403691, 0.0],
                          [3.00094461, -1.55173182, 0.0]]},
        ],
    }
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
