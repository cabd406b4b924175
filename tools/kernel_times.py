#!/usr/bin/env python
"""Per-kernel CUDA-event times (the library's own profiler) for one rasterize_rgbad fwd+bwd at a given shape.

    python tools/kernel_times.py --batch 32 --faces 70000 --size 512 --ts 2 --rgb 1 --alpha 0 --depth 1
"""
import argparse
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neural_renderer as nr  # noqa: E402
from neural_renderer_b200 import _lib, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--faces", type=int, default=5000)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--ts", type=int, default=4)
    ap.add_argument("--aa", type=int, default=0)
    ap.add_argument("--rgb", type=int, default=1)
    ap.add_argument("--alpha", type=int, default=0)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda")
    fa = torch.from_numpy(synthetic.sphere_faces(a.batch, a.faces)).to(dev).requires_grad_(True)
    ta = torch.from_numpy(synthetic.random_textures(a.batch, a.faces, a.ts)).to(dev).requires_grad_(True) if a.rgb else None
    keys = [k for k, on in (("rgb", a.rgb), ("alpha", a.alpha), ("depth", a.depth)) if on]

    def step():
        fa.grad = None
        if ta is not None:
            ta.grad = None
        o = nr.rasterize_rgbad(fa, ta, a.size, bool(a.aa), 0.1, 100, 1e-4, [0, 0, 0], bool(a.rgb), bool(a.alpha), bool(a.depth))
        outs = [o[k] for k in keys]
        torch.autograd.backward(outs, [torch.ones_like(t) if k != "rgb" else torch.randn_like(t) for k, t in zip(keys, outs)])

    for _ in range(3):
        step()
    lib = _lib.load()
    lib.nr_b200_set_profiling(1)
    _lib.read_profile()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    prof = _lib.read_profile()
    lib.nr_b200_set_profiling(0)
    acc = collections.OrderedDict()
    for name, ms in prof:
        acc[name] = acc.get(name, 0.0) + ms / a.steps
    print(json.dumps({"shape": vars(a), "kernels_ms_per_step": acc, "sum_ms": sum(acc.values())}))


if __name__ == "__main__":
    main()
