"""Example 4 of the reference (examples/example4.py:18-57, 93-113) with torch in place of Chainer: find the camera
position from which a mesh's silhouette matches a target image.

The call sequence is the reference's: `Renderer()`, `renderer.eye = <parameter>`, `render_silhouettes`, sum of squared
differences, Adam(0.1).  The gradient reaches the camera through the fused look_at + perspective kernel
(`nr_b200_camera_transform_backward`: d loss / d eye through the translation and through the look_at rotation).  The
target is rendered from a known eye instead of being read from examples/data/example4_ref.png.

    python examples/example4_optimize_camera.py [--iters 200]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import neural_renderer  # noqa: E402


class Model(torch.nn.Module):
    def __init__(self, vertices, faces, image_ref, start):
        super().__init__()
        self.register_buffer("vertices", vertices[None, :, :])
        self.register_buffer("faces", faces[None, :, :])
        self.register_buffer("image_ref", image_ref)
        self.camera_position = torch.nn.Parameter(torch.tensor(start, dtype=torch.float32))
        self.renderer = neural_renderer.Renderer()
        self.renderer.eye = self.camera_position

    def forward(self):
        image = self.renderer.render_silhouettes(self.vertices, self.faces)
        return ((image - self.image_ref[None, :, :]) ** 2).sum()


def run(iters=200, device="cuda", start=(6.0, 10.0, -14.0)):
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "teapot.npz"))
    vertices, faces = torch.from_numpy(d["vertices"]).to(device), torch.from_numpy(d["faces"]).to(device)
    with torch.no_grad():
        r = neural_renderer.Renderer()
        r.eye = neural_renderer.get_points_from_angles(2.732, 30, -15)
        target = r.render_silhouettes(vertices[None], faces[None])[0]
    model = Model(vertices, faces, target, start).to(device)
    optimizer = neural_renderer.Adam(model.parameters(), lr=0.1)
    losses = []
    for _ in range(iters):
        optimizer.zero_grad()
        loss = model()
        loss.backward()
        optimizer.step()
        losses.append(float(loss.detach()))
        if losses[-1] < 70:  # the reference's stopping rule
            break
    return losses, model.camera_position.detach().cpu().numpy()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    a = ap.parse_args()
    losses, eye = run(a.iters)
    print("loss %.1f -> %.1f in %d iterations; camera at %s" % (losses[0], losses[-1], len(losses), np.round(eye, 3)))
