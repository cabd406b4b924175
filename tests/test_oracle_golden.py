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


def test_row_coverage_is_one_interval_in_fp32():
    """The forward kernel's row-span rasterization rests on this: for a fixed pixel row every edge test of the
    reference, (yp - y_k) * dx_k < (xp - x_k) * dy_k evaluated in fp32 (sub, sub, mul -- never fused), flips at most
    once along x, so the pixels that pass all three tests form ONE interval.  Checked here by brute force over random
    and adversarial triangles (tiny, huge, nearly degenerate, vertices on pixel centres) with numpy's IEEE float32."""
    rng = np.random.default_rng(2024)
    f32 = np.float32
    for S in (17, 64, 256):
        xs = ((2 * np.arange(S) + 1 - S).astype(np.float64) / S).astype(f32)          # rasterize.py:291-292
        n = 1500
        scale = rng.choice([1e-3, 3e-2, 0.3, 1.0, 5.0, 50.0], size=(n, 1, 1))
        tri = (rng.uniform(-1, 1, size=(n, 1, 2)) + scale * rng.uniform(-1, 1, size=(n, 3, 2))).astype(f32)
        tri[:50, :, 0] = xs[rng.integers(0, S, size=(50, 3))]                          # vertices exactly on pixel centres
        tri[50:100, :, 1] = xs[rng.integers(0, S, size=(50, 3))]
        tri[100:130, 2] = tri[100:130, 1] + f32(1e-6) * (tri[100:130, 0] - tri[100:130, 1])   # slivers
        for t in tri:
            x0, y0, x1, y1, x2, y2 = t[0, 0], t[0, 1], t[1, 0], t[1, 1], t[2, 0], t[2, 1]
            yp = xs[:, None]                                                            # rows
            xp = xs[None, :]                                                            # columns
            edges = ((x0, y0, x1 - x0, y1 - y0), (x1, y1, x2 - x1, y2 - y1), (x2, y2, x0 - x2, y0 - y2))
            inside = np.ones((S, S), dtype=bool)
            for (xe, ye, dx, dy) in edges:
                out = ((yp - ye) * dx) < ((xp - xe) * dy)                               # [rows, cols], fp32 throughout
                flips = (out[:, 1:] != out[:, :-1]).sum(axis=1)
                assert flips.max() <= 1                                                 # monotone along x
                if flips.any():
                    # the direction of the flip is the sign of dy, as the kernel's binary search assumes
                    r = int(np.argmax(flips))
                    assert out[r, -1] == (dy >= 0)
                inside &= ~out
            runs = (inside[:, 1:] != inside[:, :-1]).sum(axis=1) + inside[:, 0] + inside[:, -1]
            assert runs.max() <= 2                                                      # covered pixels: one interval per row


def test_needles_win_pixels_beyond_their_vertex_box_and_the_margin_covers_them():
    """What the forward's thin-face margin (csrc/nr_bbox.cuh: thin_face_margin) is for, on the CPU oracle: needles whose
    long edges meet at 1e-7 .. 1e-4 rad pass the reference's fp32 edge tests (rasterize.py:309-311) at a pixel centre
    beyond their tip and WIN it; the margin formula, replayed in numpy, reaches every such pixel."""
    import nr_oracle as o
    from neural_renderer_b200 import synthetic
    S, F = 64, 100
    faces = synthetic.needle_faces(1, F, S, seed=3)
    r = o.OracleRasterize(S, 0.1, 100, 1e-4, (0, 0, 0), return_alpha=True)
    r.forward(faces)
    fim = r.face_index_map[0]
    ys, xs = np.nonzero(fim >= 0)
    v = faces[0, fim[ys, xs]]
    px, py = 0.5 * (v[:, :, 0] * S + S - 1), 0.5 * (v[:, :, 1] * S + S - 1)
    over = np.maximum.reduce([xs - px.max(1), px.min(1) - xs, ys - py.max(1), py.min(1) - ys])  # pixels beyond the vertices
    outside = over > 1.0 / 256
    assert outside.sum() >= F // 2
    e = np.stack([v[:, 1, :2] - v[:, 0, :2], v[:, 2, :2] - v[:, 1, :2], v[:, 0, :2] - v[:, 2, :2]], 1).astype(np.float32)
    l2 = (e ** 2).sum(-1).max(1)
    det = np.abs(e[:, 0, 0] * e[:, 1, 1] - e[:, 0, 1] * e[:, 1, 0])
    D = 1.4143 * (1 + np.abs(v[:, :, :2]).max((1, 2)))
    with np.errstate(divide="ignore", invalid="ignore"):
        ext = 16 * 2.0 ** -24 * D * S * l2 / det
    margin = 1.0 / 256 + np.where(ext < 8.0, ext, 8.0)  # NaN / inf -> the cap, like the kernel
    assert (over[outside] <= 0.5 * margin[outside]).all()
