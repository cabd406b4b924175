"""Example 3 of the reference (examples/example3.py:18-53) with torch in place of Chainer: optimise per-face textures
so that renders from random viewpoints match a target image (here: a uniform colour, so the script is self-contained).

    python examples/example3_optimize_textures.py [--iters 50]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import neural_renderer  # noqa: E402


def run(iters=50, device="cuda", seed=0):
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "teapot.npz"))
    vertices = torch.from_numpy(d["vertices"]).to(device)[None]
    faces = torch.from_numpy(d["faces"]).to(device)[None]
    textures = torch.zeros((1, faces.shape[1], 4, 4, 4, 3), device=device, requires_grad=True)
    target = torch.tensor([0.8, 0.3, 0.1], device=device)[None, :, None, None]
    renderer = neural_renderer.Renderer()
    renderer.perspective = False
    renderer.light_intensity_directional = 0.0
    renderer.light_intensity_ambient = 1.0
    optimizer = neural_renderer.Adam([textures], lr=0.1, betas=(0.5, 0.999))
    rng = np.random.default_rng(seed)
    losses = []
    for _ in range(iters):
        renderer.eye = neural_renderer.get_points_from_angles(2.732, 0, float(rng.uniform(0, 360)))
        optimizer.zero_grad()
        image = renderer.render(vertices, faces, torch.tanh(textures))
        mask = renderer.render_silhouettes(vertices, faces).detach()[:, None]
        loss = (((image - target) * mask) ** 2).sum()
        loss.backward()
        optimizer.step()
        losses.append(float(loss.detach()))
    return losses


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ls = run(ap.parse_args().iters)
    print("loss: first %.1f -> last %.1f" % (ls[0], ls[-1]))
