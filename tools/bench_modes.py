#!/usr/bin/env python
"""forward-only and forward+backward, separately for silhouette / RGB / depth, at the headline
shape (B=64, F=5000, 256x256, anti-aliasing off) on one GPU, each against its own algorithmic-byte roofline; then
BASELINE.json configs[2] (70k-face mesh, depth + RGB, 512x512, batch 32) with a face_index_map parity check against
the re-hosted reference kernels on two items.  CUDA events, median of N.  Prints one JSON object.

    python tools/bench_modes.py > profiles/r01c_modes.json
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
from neural_renderer_b200 import synthetic  # noqa: E402
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


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        for k in ("hbm_gbs", "hbm_copy_gbs", "hbm_gbps"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 7700.0, "nominal fallback"


def algorithmic_bytes(mode, B, F, S, ts):
    """SURVEY.md 8(d) formulas (forward, backward)."""
    P, T = B * S * S, ts ** 3
    if mode == "silhouette":
        return 36 * B * F + 8 * P, 72 * B * F + 12 * P
    if mode == "depth":
        # backward of depth: faces r/w, depth, fim, weight map, upstream gradient (the survey lists no formula)
        return 36 * B * F + 20 * P, 72 * B * F + 24 * P
    return 36 * B * F + 12 * T * B * F + 32 * P, 72 * B * F + 12 * T * B * F + 40 * P


def main():
    dev = torch.device("cuda")
    peak, peak_src = peak_gbs()
    out = {"device": torch.cuda.get_device_name(0), "timing": "CUDA events, median of 20 after 3 warm-ups",
           "hbm_peak_gbs": peak, "peak_source": peak_src, "headline_modes": [], "config3": None}
    B, F, S, ts = 64, 5000, 256, 4
    faces = torch.from_numpy(synthetic.sphere_faces(B, F)).to(dev)
    tex = torch.from_numpy(synthetic.random_textures(B, F, ts)).to(dev)
    gen = torch.Generator(device="cpu").manual_seed(99)
    g3 = torch.randn((B, 3, S, S), generator=gen).to(dev)
    g1 = torch.randn((B, S, S), generator=gen).to(dev)
    fa = faces.clone().requires_grad_(True)
    ta = tex.clone().requires_grad_(True)
    calls = {
        "silhouette": (lambda: nr.rasterize_silhouettes(fa, S, False), g1),
        "rgb": (lambda: nr.rasterize(fa, ta, S, False), g3),
        "depth": (lambda: nr.rasterize_depth(fa, S, False), g1),
    }
    for mode, (fwd, g) in calls.items():
        def fb():
            fa.grad = None
            ta.grad = None
            fwd().backward(g)
        with torch.no_grad():
            t_f = timeit(fwd)
        t_fb = timeit(fb)
        bf, bb = algorithmic_bytes(mode, B, F, S, ts)
        out["headline_modes"].append({
            "mode": mode, "fwd_ms": t_f, "fwd_bwd_ms": t_fb,
            "fwd_mpixels_per_s": B * S * S / t_f / 1e3, "fwd_bwd_mpixels_per_s": B * S * S / t_fb / 1e3,
            "fwd_algorithmic_mb": bf / 1e6, "fwd_roofline_frac": bf / (t_f * 1e-3) / 1e9 / peak,
            "bwd_algorithmic_mb": bb / 1e6, "fwd_bwd_roofline_frac": (bf + bb) / (t_fb * 1e-3) / 1e9 / peak,
            "note": "times are whole API calls (torch allocation + all kernels of the pass), not single kernels"})
    del fa, ta, faces, tex, g3, g1
    torch.cuda.empty_cache()

    # BASELINE.json configs[2]: ~70k faces, depth + RGB, 512x512, batch 32 (synthetic 70k-face sphere, ts=2)
    B, F, S, ts = 32, 70000, 512, 2
    faces = torch.from_numpy(synthetic.sphere_faces(B, F)).to(dev)
    tex = torch.from_numpy(synthetic.random_textures(B, F, ts)).to(dev)
    fa = faces.clone().requires_grad_(True)
    ta = tex.clone().requires_grad_(True)
    g3 = torch.randn((B, 3, S, S), device=dev)
    g1 = torch.randn((B, S, S), device=dev)

    def fwd3():
        return nr.rasterize_rgbad(fa, ta, S, False, 0.1, 100, 1e-4, [0, 0, 0], True, False, True)

    def fb3():
        fa.grad = None
        ta.grad = None
        o = fwd3()
        torch.autograd.backward([o["rgb"], o["depth"]], [g3, g1])
    with torch.no_grad():
        t_f = timeit(fwd3, n=10)
    t_fb = timeit(fb3, n=10)
    row = {"config": "configs[2]: 70k-face sphere, depth + RGB, 512x512, batch 32, ts 2, anti-aliasing off",
           "fwd_ms": t_f, "fwd_bwd_ms": t_fb, "fwd_bwd_mpixels_per_s": B * S * S / t_fb / 1e3}
    if refhost.available(512, 70000, 2, 0.1, 100, 1e-4, 1, 0, 1):
        nb = 2
        ref = refhost.rasterize_rgbad(faces[:nb].contiguous(), tex[:nb].contiguous(), S, False, 0.1, 100, 1e-4,
                                      [0, 0, 0], True, False, True)
        t_ref = timeit(lambda: refhost.rasterize_rgbad(faces[:nb].contiguous(), tex[:nb].contiguous(), S, False, 0.1,
                                                       100, 1e-4, [0, 0, 0], True, False, True), n=3, warm=1)
        with torch.no_grad():
            got = nr.rasterize_rgbad(faces[:nb].contiguous(), tex[:nb].contiguous(), S, False, 0.1, 100, 1e-4,
                                     [0, 0, 0], True, False, True)
        row.update({
            "ref_fwd_ms_2_items": t_ref, "ref_fwd_ms_extrapolated_batch_32": t_ref * B / nb,
            "rgb_max_abs_diff_vs_reference": float((got["rgb"] - ref["rgb"]).abs().max()),
            "depth_max_abs_diff_vs_reference": float((got["depth"] - ref["depth"]).abs().max()),
            "alpha_equal_via_depth_cover": bool(torch.equal(got["depth"] < 100, ref["depth"] < 100))})
    out["config3"] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
