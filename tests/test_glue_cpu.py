"""CPU: the torch glue (look_at / perspective / lighting / vertices_to_faces) against the oracle's numpy restatement,
and a 2-process gloo check of the batch sharding used by bench.py."""
import os
import sys

import numpy as np
import pytest
import torch

import nr_oracle as o
import neural_renderer as nr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_camera_and_gather_match_oracle(teapot):
    v, f = teapot
    vb = np.stack([v, v * 0.9])
    fb = np.stack([f, f])
    eye = [1.0, 1.0, -2.7]
    ref = o.vertices_to_faces(o.perspective(o.look_at(vb, eye)), fb)
    got = nr.vertices_to_faces(nr.perspective(nr.look_at(torch.from_numpy(vb), eye)), torch.from_numpy(fb))
    np.testing.assert_allclose(got.numpy(), ref, rtol=2e-5, atol=2e-6)


def test_lighting_matches_oracle(teapot):
    v, f = teapot
    vb, fb = v[None], f[None]
    faces = o.vertices_to_faces(vb, fb)
    tex = np.random.default_rng(0).random((1, f.shape[0], 2, 2, 2, 3), dtype=np.float32)
    ref = o.lighting(faces, tex)
    got = nr.lighting(torch.from_numpy(faces), torch.from_numpy(tex))
    np.testing.assert_allclose(got.numpy(), ref, rtol=2e-5, atol=2e-6)


def test_look_and_points_from_angles():
    # tests/test_look_at.py:13-28 of the reference: eyes on the axes map the origin to z = distance
    v = torch.tensor([[[1.0, 0.0, 0.0], [0.0, 0.0, 0.0]]])
    out = nr.look_at(v, [0.0, 0.0, -2.0])
    np.testing.assert_allclose(out[0, 1].numpy(), [0, 0, 2], atol=1e-4)
    e = nr.get_points_from_angles(2.0, 0.0, 90.0)
    np.testing.assert_allclose(e, (2.0, 0.0, 0.0), atol=1e-6)
    out = nr.look(v, [0.0, 0.0, -2.0], [0.0, 0.0, 1.0])
    np.testing.assert_allclose(out[0, 0].numpy(), [1, 0, 2], atol=1e-4)


def _shard_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = bench.shard_range(10, rank, world)
    # shared-mesh gradient reduction used by config 5: sum over viewpoint shards
    from neural_renderer_b200.distributed import allreduce_shared_grads
    param = torch.zeros(4, requires_grad=True)
    param.grad = torch.full((4,), float(hi - lo))
    for w in allreduce_shared_grads([param], async_op=True):
        w.wait()
    g = param.grad
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out.put((rank, lo, hi, g.tolist(), float(t)))
    dist.destroy_process_group()


def test_batch_sharding_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 10)]  # contiguous, disjoint, complete
    assert res[0][3] == [10.0] * 4 and res[1][3] == [10.0] * 4
    assert res[0][4] == 2.0  # max over ranks (timing reduction)


def test_camera_transform_composes_look_at_and_perspective():
    """functional.camera_transform (what Renderer._transform calls) on CPU tensors == look_at / look followed by
    perspective, and differentiates with respect to a tensor eye (examples/example4 call sequence)."""
    import torch
    from neural_renderer_b200 import functional as F
    gen = torch.Generator().manual_seed(3)
    v = torch.rand((2, 40, 3), generator=gen, dtype=torch.float64) - 0.5
    eye = [0.4, 0.9, -2.5]
    assert torch.allclose(F.camera_transform(v, eye), F.perspective(F.look_at(v, eye), 30.))
    assert torch.allclose(F.camera_transform(v, eye, "look", [0.1, 0.0, 1.0], False), F.look(v, eye, [0.1, 0.0, 1.0]))
    assert torch.equal(F.camera_transform(v, eye, "none", None, False), v)
    e = torch.tensor(eye, dtype=torch.float64, requires_grad=True)
    F.camera_transform(v, e).sum().backward()
    assert e.grad is not None and torch.isfinite(e.grad).all() and float(e.grad.abs().sum()) > 0


def test_face_light_is_the_factor_of_lighting(teapot):
    """lighting(faces, textures) == textures * face_light(faces) (lighting.py:29-52), and
    face_light_from_vertices(vertices, faces) == face_light(vertices_to_faces(vertices, faces)) on the CPU path."""
    import torch
    from neural_renderer_b200 import functional as F
    v, f = teapot
    vertices = torch.from_numpy(v)[None]
    faces = torch.from_numpy(f)[None]
    tex = torch.rand((1, f.shape[0], 2, 2, 2, 3), generator=torch.Generator().manual_seed(4))
    args = (0.4, 0.6, [1.0, 0.9, 0.8], [0.3, 1.0, 0.5], [0.0, 0.8, -0.6])
    fw = F.vertices_to_faces(vertices, faces)
    light = F.face_light(fw, *args)
    assert light.shape == (1, f.shape[0], 3) and float(light.min()) >= 0.4 * 0.8 - 1e-6
    assert torch.equal(F.lighting(fw, tex, *args), tex * light[:, :, None, None, None, :])
    assert torch.equal(F.face_light_from_vertices(vertices, faces, *args), light)


def test_obj_roundtrip(tmp_path, teapot):
    from neural_renderer_b200 import io
    v, f = teapot
    path = str(tmp_path / "t.obj")
    io.save_obj(path, v, f)
    v2, f2 = io.load_obj(path, normalization=False)
    assert np.array_equal(f2, f) and np.allclose(v2, v, atol=1e-6)


def test_reference_known_answers_look_at_perspective():
    """The reference's own known-answer cases: tests/test_look_at.py:12-28 (three eyes) and
    tests/test_perspective.py:12-18 (default 30 degree viewing angle, pi = 3.1416 as in perspective.py)."""
    import torch
    v = torch.tensor([[[1.0, 0.0, 0.0]]])
    s2 = np.sqrt(2.0)
    for eye, answer in (([1, 0, 1], [-s2 / 2, 0, s2 / 2]), ([0, 0, -10], [1, 0, 10]), ([-1, 1, 0], [0, s2 / 2, 1.5 * s2])):
        out = nr.look_at(v, torch.tensor(eye, dtype=torch.float32))
        np.testing.assert_allclose(out.numpy().ravel(), answer, rtol=1e-4, atol=1e-5)   # chainer.testing default rtol
        out = nr.look_at(v, eye)                                                        # plain-number eye
        np.testing.assert_allclose(out.numpy().ravel(), answer, rtol=1e-4, atol=1e-5)
    out = nr.perspective(torch.tensor([[[1.0, 2.0, 10.0]]]))
    np.testing.assert_allclose(out.numpy().ravel(), [np.sqrt(3) / 10, 2 * np.sqrt(3) / 10, 10], rtol=1e-4, atol=1e-5)


def test_reference_known_answers_load_obj(tmp_path, teapot):
    """tests/test_load_obj.py:11-37: the tetrahedron (raw and normalised) and the teapot's element counts."""
    from neural_renderer_b200 import io
    path = tmp_path / "tetrahedron.obj"
    path.write_text("v 1 0 0\nv 0 1 0\nv 0 0 1\nv 0 0 0\nf 2 4 3\nf 4 2 1\nf 3 1 2\nf 1 3 4\n")
    v_ref = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0]], dtype=np.float32)
    f_ref = np.array([[1, 3, 2], [3, 1, 0], [2, 0, 1], [0, 2, 3]], dtype=np.int32)
    v, f = io.load_obj(str(path), False)
    assert np.allclose(v, v_ref) and np.array_equal(f, f_ref)
    v, f = io.load_obj(str(path), True)
    assert np.allclose(v, v_ref * 2 - 1.0) and np.array_equal(f, f_ref)
    tv, tf = teapot
    assert tf.shape[0] == 2464 and tv.shape[0] == 1292


def test_save_obj_with_textures_round_trip(tmp_path, monkeypatch, teapot):
    """save_obj(filename, vertices, faces, textures) (save_obj.py:10-191: atlas PNG + .mtl + vt / usemtl / f v/vt) and
    back through load_obj(load_texture=True): geometry exact, per-face textures within the atlas' 16-pixel tiles'
    resampling error (smooth textures: a few 8-bit levels)."""
    import nr_oracle as o
    from neural_renderer_b200 import io
    monkeypatch.setattr(io, "bake_textures", lambda image, uv, upd, ts, tex: o.bake_textures(image, uv, upd, ts, tex))
    v, f = teapot
    f = f[:300]
    ts = 4
    # a smooth per-face texture: colour = barycentric position inside the cube, scaled per face
    g = np.linspace(0, 1, ts, dtype=np.float32)
    cube = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1)                       # [ts,ts,ts,3]
    scale = np.random.default_rng(0).uniform(0.3, 1.0, size=(f.shape[0], 1, 1, 1, 3)).astype(np.float32)
    textures = cube[None] * scale
    path = str(tmp_path / "m.obj")
    io.save_obj(path, v, f, textures)
    for ext in (".obj", ".mtl", ".png"):
        assert os.path.exists(path[:-4] + ext)
    text = open(path).read()
    assert text.startswith("# m.obj\n#\n\nmtllib m.mtl\n") and "usemtl material_1" in text
    assert open(path[:-4] + ".mtl").read() == "newmtl material_1\nmap_Kd m.png\n"
    v2, f2, t2 = io.load_obj(path, normalization=False, texture_size=ts, load_texture=True)
    assert np.array_equal(f2, f) and np.allclose(v2, v, atol=1e-6)
    assert t2.shape == textures.shape
    # texels on the barycentric plane a + b + c = ts - 1 are the ones a face ever samples (weights sum to 1)
    a, b, c = np.meshgrid(np.arange(ts), np.arange(ts), np.arange(ts), indexing="ij")
    plane = (a + b + c) == ts - 1
    err = np.abs(t2[:, plane] - textures[:, plane])
    assert float(np.nanmax(err)) < 0.12 and float(np.nanmean(err)) < 0.03


def test_texture_atlas_layout():
    from neural_renderer_b200 import io
    tex = np.zeros((5, 2, 2, 2, 3), np.float32)
    tex[3] = 1.0
    image, uv = io.create_texture_image(tex, texture_size_out=8)
    assert image.shape == (16, 24, 3) and uv.shape == (5, 3, 2)      # 3 x 2 tiles of 8 pixels
    assert uv.min() >= 0 and uv.max() <= 1
    up = image[::-1]                                                   # un-flip: tile (row 1, column 0) is face 3
    assert abs(up[8:16, 0:8].max() - 1.0) < 1e-5 and up[0:8].max() == 0.0 and up[8:16, 8:].max() == 0.0


def test_adam_masks_zero_gradients_and_honours_param_lr():
    """optimizers.py:9-39: elements with grad == 0 keep parameter and both moments; lr = self.lr * param.lr."""
    import torch
    from neural_renderer_b200.optimizers import Adam
    p = torch.nn.Parameter(torch.tensor([1.0, 2.0, 3.0, 4.0]))
    q = torch.nn.Parameter(torch.tensor([1.0, 2.0]))
    q.lr = 0.0
    opt = Adam([p, q], lr=0.1, betas=(0.9, 0.999))
    p.grad = torch.tensor([0.5, 0.0, -0.25, 0.0])
    q.grad = torch.tensor([1.0, 1.0])
    opt.step()
    # first Adam step moves by lr * sign(g) (bias-corrected), untouched where g == 0; q.lr = 0 freezes q
    np.testing.assert_allclose(p.detach().numpy(), [0.9, 2.0, 3.1, 4.0], rtol=1e-5)
    assert torch.equal(q.detach(), torch.tensor([1.0, 2.0]))
    m = opt.state[p]["m"].clone()
    p.grad = torch.tensor([0.0, 1.0, 0.0, 0.0])
    opt.step()
    assert opt.state[p]["m"][0] == m[0] and opt.state[p]["m"][2] == m[2]        # moments frozen where grad == 0
    assert p[0].item() == pytest.approx(0.9) and p[1].item() < 2.0
    # a closure's gradients define the mask (not the stale ones from before the call)
    p.grad = torch.tensor([1.0, 1.0, 1.0, 1.0])
    before = p.detach().clone()

    def closure():
        p.grad = torch.tensor([0.0, 0.0, 0.0, 2.0])
        return torch.tensor(0.0)
    opt.step(closure)
    assert torch.equal(p.detach()[:3], before[:3]) and p[3].item() < before[3].item()


def test_mesh_get_batch_is_a_broadcast_and_set_lr(tmp_path, teapot):
    import torch
    from neural_renderer_b200 import io
    from neural_renderer_b200.mesh import Mesh
    v, f = teapot
    path = str(tmp_path / "t.obj")
    io.save_obj(path, v, f)
    mesh = Mesh(path, texture_size=2)
    vv, ff, tt = mesh.get_batch(3)
    assert vv.shape == (3, v.shape[0], 3) and ff.shape == (3, f.shape[0], 3) and tt.shape == (3, f.shape[0], 2, 2, 2, 3)
    assert vv.stride(0) == 0 and ff.stride(0) == 0 and tt.stride(0) == 0      # views, not copies (mesh.py:29-34)
    assert torch.equal(tt[1], torch.sigmoid(mesh.textures))
    tt.sum().backward()
    assert torch.allclose(mesh.textures.grad, 3 * torch.sigmoid(mesh.textures) * (1 - torch.sigmoid(mesh.textures)), atol=1e-6)
    mesh.set_lr(0.5, 2.0)
    assert mesh.vertices.lr == 0.5 and mesh.textures.lr == 2.0
