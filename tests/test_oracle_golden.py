"""CPU: pins the oracle (oracle/nr_oracle.c) against the reference's own golden vectors."""
import numpy as np
import pytest

import nr_oracle as o
from helpers import to_minibatch


@pytest.fixture(scope="module")
def teapot_batch(teapot):
    v, f = teapot
    tex = np.ones((f.shape[0], 4, 4, 4, 3), np.float32)
    return to_minibatch((v, f, tex))


def test_silhouette_matches_blender(teapot_batch, golden_images):
    # tests/test_rasterize_silhouettes.py:15-35
    v, f, _ = teapot_batch
    r = o.Renderer()
    r.image_size, r.anti_aliasing = 256, False
    img = r.render_silhouettes(v, f)["alpha"][2]
    assert (img != golden_images["silhouette"]).sum() == 0
    assert img.sum() == 7580


def test_depth_matches_fixtures(teapot_batch, golden_images):
    # tests/test_rasterize_depth.py:16-58
    v, f, _ = teapot_batch
    r = o.Renderer()
    r.image_size, r.anti_aliasing = 256, False
    im = r.render_depth(v, f)["depth"][2].copy()
    assert ((im != im.max()).astype(np.float32) != golden_images["silhouette"]).sum() == 0
    im[im == im.max()] = im.min()
    im = (im - im.min()) / (im.max() - im.min())
    np.testing.assert_allclose(im, golden_images["depth_u8"].astype(np.float32) / 255., atol=1e-2)


def test_rgb_matches_blender(teapot_batch, golden_images):
    # tests/test_rasterize.py:52-74
    v, f, tex = teapot_batch
    r = o.Renderer()
    r.image_size, r.anti_aliasing = 256, False
    r.light_intensity_ambient, r.light_intensity_directional = 1.0, 0.0
    img = r.render(v, f, tex)["rgb"][2].mean(0)
    np.testing.assert_allclose(img, golden_images["silhouette"], rtol=1e-4, atol=1e-5)


def test_soft_golden_snapshots(teapot_batch, golden_images):
    # tests/test_rasterize.py:15-50 write these PNGs without asserting; 8-bit, min-max scaled by scipy.misc.imsave
    v, f, tex = teapot_batch
    r = o.Renderer()
    r.image_size, r.anti_aliasing = 256, False
    img = r.render(v, f, tex)["rgb"][2].transpose(1, 2, 0)
    img = (img - img.min()) / (img.max() - img.min())
    ref = golden_images["rasterize1_u8"].astype(np.float32) / 255.
    assert (np.abs(img - ref) > 2.5 / 255).sum() <= 64  # lighting / flip conventions; a few edge pixels differ
    r = o.Renderer()
    r.eye = [1, 1, -2.7]
    img = r.render(v, f, tex)["rgb"][2].transpose(1, 2, 0)
    img = (img - img.min()) / (img.max() - img.min())
    ref = golden_images["rasterize2_u8"].astype(np.float32) / 255.
    assert (np.abs(img - ref) > 2.5 / 255).sum() <= 256  # anti-aliasing conventions


@pytest.mark.parametrize("mode", ["silhouette", "rgb"])
def test_known_answer_gradients(kat, mode):
    # tests/test_rasterize_silhouettes.py:37-99, tests/test_rasterize.py:76-149 (rtol 1e-2 there)
    for c in kat["cases"]:
        r = o.Renderer()
        r.image_size, r.anti_aliasing, r.perspective = 64, False, False
        vv, ff, gref = to_minibatch((np.array(c["vertices"], np.float32), np.array(c["faces"], np.int32),
                                     np.array(c["grad_ref"], np.float32)))
        minus = 1.0 if c["name"] == "out_of_face" else 0.0
        if mode == "silhouette":
            res = r.render_silhouettes(vv, ff)
            img = res["alpha"]
            g = np.zeros_like(img)
            g[:, c["pyi"], c["pxi"]] = np.sign(img[:, c["pyi"], c["pxi"]] - minus)
            gv, _ = res.backward_vertices(grad_alpha=g)
            tol = 1e-4  # the reference vectors were produced with eps = 1e-4, which this path uses
        else:
            r.light_intensity_ambient, r.light_intensity_directional = 1.0, 0.0
            tt, = to_minibatch((np.ones((1, 4, 4, 4, 3), np.float32),))
            res = r.render(vv, ff, tt)
            img = res["rgb"].mean(1)
            gi = np.zeros_like(img)
            gi[:, c["pyi"], c["pxi"]] = np.sign(img[:, c["pyi"], c["pxi"]] - minus)
            gv, _ = res.backward_vertices(grad_rgb=np.repeat(gi[:, None], 3, axis=1) / 3)
            tol = 1e-2  # Renderer.render passes eps = 1e-3 (renderer.py:105)
        np.testing.assert_allclose(gv, gref, rtol=tol, atol=tol * 1e-2)


def test_depth_gradient_matches_finite_differences():
    # the analytic K7 term (rasterize.py:805-847) against central differences (the reference's own check,
    # tests/test_rasterize_depth.py:60-93, indexes an empty batch slot and is vacuous)
    verts = np.array([[-0.9, -0.9, 2.], [-0.8, 0.8, 1.], [0.8, 0.8, 0.5]], np.float32)
    faces = verts[None, None, :, :]  # [1,1,3,3], already front-facing in the rasterizer's convention?
    if o.rasterize_depth(faces, 64, False)["depth"].min() >= 100:
        faces = faces[:, :, ::-1].copy()
    py, px = 15, 20
    res = o.rasterize_depth(faces, 64, False)
    d = res["depth"]
    assert d[0, py, px] < 100
    g = np.zeros_like(d)
    g[0, py, px] = 2 * (d[0, py, px] - 1)
    gf, _ = res.backward(grad_depth=g)
    num = np.zeros_like(gf)
    h = 1e-3
    for k in range(3):
        for l in range(3):
            fp, fm = faces.copy(), faces.copy()
            fp[0, 0, k, l] += h
            fm[0, 0, k, l] -= h
            lp = (o.rasterize_depth(fp, 64, False)["depth"][0, py, px] - 1) ** 2
            lm = (o.rasterize_depth(fm, 64, False)["depth"][0, py, px] - 1) ** 2
            num[0, 0, k, l] = (lp - lm) / (2 * h)
    np.testing.assert_allclose(gf, num, atol=2e-3)
