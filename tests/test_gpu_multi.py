"""2-GPU test of the one collective the path has (BASELINE.json configs[4]): a mesh shared by all
viewpoints, viewpoints sharded over the ranks, vertex / texture gradients summed with NCCL -- the texture all-reduce
launched between the two halves of the rasterizer's backward (overlap_texture_allreduce).  N-rank reduced gradients
must equal the gradients one rank computes over ALL viewpoints.  Skipped with fewer than 2 devices
(run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _render_views(nr, vertices, textures, faces_idx, eyes, grad, S, fill_back):
    r = nr.Renderer()
    r.image_size, r.anti_aliasing, r.fill_back = S, False, fill_back
    r.eye = eyes
    # rasterize.py:389 (the sampler reads the vertex depths of batch item 0 OF THE CALL) makes the reference's result depend
    # on how the viewpoints are batched; sharding changes which viewpoint is "item 0".  The sharded and the single-rank
    # run can only be compared with every item sampling with its own depths.
    r.reference_exact = False
    V = eyes.shape[0]
    img = r.render(vertices[None].expand(V, -1, -1), faces_idx[None].expand(V, -1, -1), textures[None])
    (img * grad).sum().backward()
    return img.detach()


def _worker(rank, world, port, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    import neural_renderer_b200 as nr
    from neural_renderer_b200 import synthetic
    from neural_renderer_b200.distributed import allreduce_shared_grads, overlap_texture_allreduce, shard_range
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    F, S, ts, V_total = 20000, 256, 2, 6
    v_np, f_np = synthetic.sphere_mesh(F)
    faces_idx = torch.from_numpy(f_np).to(dev)
    az = torch.arange(V_total, dtype=torch.float32) * (360.0 / V_total)
    eyes_all = nr.get_points_from_angles(torch.full_like(az, 2.732), torch.full_like(az, 25.0), az).to(dev)
    grad_all = torch.randn((V_total, 3, S, S), generator=torch.Generator().manual_seed(5)).to(dev)
    tex0 = torch.rand((F, ts, ts, ts, 3), generator=torch.Generator().manual_seed(7))
    result = {}
    for fill_back in (False, True):
        vertices = torch.from_numpy((v_np * 0.55).astype(np.float32)).to(dev).requires_grad_(True)
        textures = tex0.to(dev).requires_grad_(True)
        lo, hi = shard_range(V_total, rank, world)
        with overlap_texture_allreduce() as ov:
            _render_views(nr, vertices, textures, faces_idx, eyes_all[lo:hi], grad_all[lo:hi], S, fill_back)
        allreduce_shared_grads([vertices])
        torch.cuda.synchronize()
        assert ov.launched == 1
        if rank == 0:
            v1 = vertices.detach().clone().requires_grad_(True)
            t1 = textures.detach().clone().requires_grad_(True)
            _render_views(nr, v1, t1, faces_idx, eyes_all, grad_all, S, fill_back)  # every viewpoint on one rank
            torch.cuda.synchronize()

            def rel(a, b):
                return float((a - b).abs().max() / b.abs().max())
            result["fill_back=%s" % fill_back] = (rel(vertices.grad, v1.grad), rel(textures.grad, t1.grad))
    dist.barrier(device_ids=[rank])
    if rank == 0:
        torch.save(result, out_path)
    dist.destroy_process_group()


def test_two_rank_shared_mesh_gradients_equal_single_rank(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    result = torch.load(out)
    assert set(result) == {"fill_back=False", "fill_back=True"}
    for k, (ev, et) in result.items():
        assert ev <= 1e-5 and et <= 1e-5, (k, ev, et)
