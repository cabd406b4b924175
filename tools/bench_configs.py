#!/usr/bin/env python
"""Time the smaller BASELINE.json configurations (teapot through `Renderer`) on one GPU: ours vs the re-hosted
reference kernels, forward and forward+backward, CUDA events, median of N.  Prints one JSON object.

    python tools/bench_configs.py > profiles/r01_configs.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import neural_renderer as nr  # noqa: E402
import refhost  # noqa: E402


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def teapot_inputs(B, dev, eye):
    d = np.load(os.path.join(ROOT, "tests", "golden", "teapot.npz"))
    v = torch.from_numpy(np.stack([d["vertices"]] * B)).to(dev)
    f = torch.from_numpy(np.stack([d["faces"]] * B)).to(dev)
    fi = torch.cat((f, f.flip(2)), dim=1)
    faces = nr.vertices_to_faces(nr.perspective(nr.look_at(v, eye)), fi).contiguous()
    return v, f, fi, faces


def main():
    dev = torch.device("cuda")
    out = {"device": torch.cuda.get_device_name(0), "timing": "CUDA events, median of 20 after 3 warm-ups", "configs": []}
    eye = nr.get_points_from_angles(2.732, 30, 40)

    # configs[0]: teapot silhouette 64x64 (anti-aliased -> raster 128), batch 1
    v, f, fi, faces = teapot_inputs(1, dev, eye)
    g = torch.randn((1, 64, 64), device=dev)
    fa = faces.clone().requires_grad_(True)

    def ours_fwd():
        return nr.rasterize_silhouettes(fa, 64, True)

    def ours_fb():
        fa.grad = None
        ours_fwd().backward(g)

    row = {"config": "configs[0]: teapot (4928 faces with fill_back) silhouette 64x64, anti-aliasing, batch 1",
           "ours_fwd_ms": timeit(ours_fwd), "ours_fwd_bwd_ms": timeit(ours_fb)}
    if refhost.available(128, 4928, 0, 0.1, 100, 1e-4, 0, 1, 0):
        def ref_fwd():
            return refhost.rasterize_rgbad(faces, None, 64, True, 0.1, 100, 1e-4, None, False, True, False)

        def ref_fb():
            ref_fwd().backward(None, g, None)
        row.update(ref_fwd_ms=timeit(ref_fwd), ref_fwd_bwd_ms=timeit(ref_fb))
    out["configs"].append(row)

    # configs[1]: teapot RGB + texture 256x256 (raster 512), batch 8, fwd + bwd (examples 2/3 shape)
    B = 8
    v, f, fi, faces = teapot_inputs(B, dev, eye)
    tex = torch.rand((B, f.shape[1], 4, 4, 4, 3), generator=torch.Generator().manual_seed(1)).to(dev)
    tx = nr.lighting(nr.vertices_to_faces(v, fi), torch.cat((tex, tex.permute(0, 1, 4, 3, 2, 5)), dim=1)).contiguous()
    g = torch.randn((B, 3, 256, 256), device=dev)
    fa = faces.clone().requires_grad_(True)
    ta = tx.clone().requires_grad_(True)

    def ours_fwd2():
        return nr.rasterize(fa, ta, 256, True, 0.1, 100, 1e-3, [0, 0, 0])

    def ours_fb2():
        fa.grad = None
        ta.grad = None
        ours_fwd2().backward(g)

    row = {"config": "configs[1]: teapot RGB + textures 256x256, anti-aliasing (raster 512), batch 8, ts 4",
           "ours_fwd_ms": timeit(ours_fwd2), "ours_fwd_bwd_ms": timeit(ours_fb2),
           "mpixels_per_s_fwd_bwd": None}
    row["mpixels_per_s_fwd_bwd"] = B * 256 * 256 / (row["ours_fwd_bwd_ms"] * 1e-3) / 1e6
    if refhost.available(512, 4928, 4, 0.1, 100, 1e-3, 1, 0, 0):
        def ref_fwd2():
            return refhost.rasterize_rgbad(faces, tx, 256, True, 0.1, 100, 1e-3, [0, 0, 0], True, False, False)

        def ref_fb2():
            ref_fwd2().backward(g, None, None)
        row.update(ref_fwd_ms=timeit(ref_fwd2, n=5), ref_fwd_bwd_ms=timeit(ref_fb2, n=5))
    out["configs"].append(row)

    # whole Renderer.render call (camera + lighting + gather + rasterize), same shape: fused glue kernels (camera,
    # face lighting, lighting / fill_back folded into the sampler) vs the op-by-op torch formulation
    vv = v.clone().requires_grad_(True)
    tt = tex.clone().requires_grad_(True)
    for fused in (True, False):
        r = nr.Renderer()
        r.eye = eye
        r.fused = fused

        def facade():
            vv.grad = None
            tt.grad = None
            r.render(vv, f, tt).backward(g)

        def facade_fwd():
            with torch.no_grad():
                r.render(vv, f, tt)

        out["configs"].append({"config": "Renderer.render end to end, teapot 256x256 AA batch 8 (%s)"
                                         % ("fused glue" if fused else "op-by-op torch glue + materialised textures"),
                               "ours_fwd_ms": timeit(facade_fwd), "ours_fwd_bwd_ms": timeit(facade)})
    # the same fused step captured in a CUDA graph (launch latency of the ~14 glue kernels and allocations removed)
    r = nr.Renderer()
    r.eye = eye

    def step():
        vv.grad = None
        tt.grad = None
        r.render(vv, f, tt).backward(g)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    out["configs"].append({"config": "Renderer.render fwd+bwd, fused glue, replayed as one CUDA graph",
                           "ours_fwd_bwd_ms": timeit(graph.replay)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
