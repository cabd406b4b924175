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
  display/              the small textured model of test_load_obj.py:55-62 (tests/data/4e49.../model.obj, 3644 faces,
                        7 materials, 2 texture images), RE-SERIALISED by this script from the parsed arrays (model.obj
                        / model.mtl / texture*.png, lossless PNG of the decoded JPEGs) plus display_u8.npz = the
                        tests/data/display.png the reference's test writes for it (never asserted there)
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
    write_display_fixture(data)
    print("wrote", sorted(os.listdir(HERE)))


def write_display_fixture(data):
    """Parse the reference's textured test model with THIS repo's loader, write it back as a compact OBJ / MTL / PNG
    trio and check that loading the copy reproduces the parsed arrays exactly."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from neural_renderer_b200 import io
    src = os.path.join(data, "4e49873292196f02574b5684eaec43e9")
    out = os.path.join(HERE, "display")
    os.makedirs(out, exist_ok=True)
    obj = os.path.join(src, "model.obj")
    v, f = io.load_obj(obj, normalization=False)
    uv, names = io.parse_texture_faces(obj)
    colors, files = io.load_mtl(os.path.join(src, "model.mtl"))
    png = {}
    for k, (mat, rel) in enumerate(files.items()):
        im = Image.open(os.path.join(src, rel)).convert("RGB")
        png[mat] = "texture%d.png" % k
        im.save(os.path.join(out, png[mat]))
    with open(os.path.join(out, "model.mtl"), "w") as fh:
        for mat, col in colors.items():
            fh.write("newmtl %s\nKd %.9g %.9g %.9g\n" % (mat, col[0], col[1], col[2]))
            if mat in png:
                fh.write("map_Kd %s\n" % png[mat])
            fh.write("\n")
    with open(os.path.join(out, "model.obj"), "w") as fh:
        fh.write("mtllib model.mtl\n")
        for p in v:
            fh.write("v %.9g %.9g %.9g\n" % tuple(p))
        flat = uv.reshape(-1, 2)
        for t in flat:                      # one vt per face corner (already wrapped; values <= 1 stay as they are)
            fh.write("vt %.9g %.9g\n" % tuple(t))
        cur = None
        for i, face in enumerate(f):
            if names[i] != cur:
                cur = names[i]
                fh.write("usemtl %s\n" % cur)
            fh.write("f %d/%d %d/%d %d/%d\n" % (face[0] + 1, 3 * i + 1, face[1] + 1, 3 * i + 2, face[2] + 1, 3 * i + 3))
    v2, f2 = io.load_obj(os.path.join(out, "model.obj"), normalization=False)
    uv2, names2 = io.parse_texture_faces(os.path.join(out, "model.obj"))
    c2, files2 = io.load_mtl(os.path.join(out, "model.mtl"))
    assert np.array_equal(v2, v) and np.array_equal(f2, f) and np.array_equal(uv2, uv) and names2 == names
    assert list(c2) == list(colors) and all(np.array_equal(c2[k], colors[k]) for k in colors) and list(files2) == list(files)
    np.savez_compressed(os.path.join(out, "display_u8.npz"), display_u8=np.array(Image.open(os.path.join(data, "display.png"))))


if __name__ == "__main__":
    main()
