"""CPU: the torch glue (look_at / perspective / lighting / vertices_to_faces) against the oracle's numpy restatement,
and a 2-process gloo check of the batch sharding used by bench.py."""
import os
import sys

import numpy as np
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
