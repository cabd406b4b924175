#!/usr/bin/env python
"""bench.py -- headline benchmark of the rasterizer hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

metric   Mpixels/s, forward + backward, pixels = B x image_size^2 output pixels
workload BASELINE.json headline / SURVEY.md 8(d): B = 64 independent 5000-face meshes per GPU (seeded UV-spheres,
         neural_renderer_b200/synthetic.py), image 256 x 256, anti_aliasing off, textures [B,F,4,4,4,3] ~ U(0,1),
         `rasterize(faces, textures)` forward + backward with a dense N(0,1) upstream gradient (RGB path:
         grad wrt faces and textures).  N > 1: one process per GPU, every rank renders its own 64 meshes (batch
         sharding, no data-path collective) -> "scaling": "weak"; value = all ranks' pixels / max-over-ranks time.
step     one forward + backward pass over one batch.

One JSON line on rank 0.  Besides the base contract it carries
  roofline      dominant kernel (by device time) against the measured HBM peak (MEASURED_PEAKS.json),
  roofline_fwd  the same for the forward raster kernel (the north-star's >= 70 % target is on forward rasterize),
  cpu_baseline  the CPU oracle (oracle/nr_oracle.c, a port -- the reference ships no CPU path) on a bounded sample,
  reference_gpu the reference's own CuPy kernels re-hosted (oracle/_ref) on the same GPU, same inputs,
  kernels       average device time per kernel of one step (CUDA events on the launching stream, separate pass).
`--impl reference` runs the reference's own implementation of the path: its unmodified CUDA kernel strings
re-hosted without CuPy (oracle/refhost.py + oracle/_ref/*.so) on the GPU -- the reference has no CPU path; if those
binaries are missing, the CPU oracle port is timed instead (and the line says so).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOAD = dict(batch_per_gpu=64, num_faces=5000, image_size=256, texture_size=4, anti_aliasing=False,
                near=0.1, far=100, eps=1e-4, background=(0.0, 0.0, 0.0))


from neural_renderer_b200.distributed import shard_range  # noqa: E402,F401  (re-exported for the tests)


def algorithmic_bytes(B, F, S, ts):
    """compulsory traffic of the RGB passes (inputs once, outputs + saved maps once)."""
    P, T = B * S * S, ts ** 3
    fwd = 36 * B * F + 12 * T * B * F + 32 * P
    bwd = 72 * B * F + 12 * T * B * F + 40 * P
    return fwd, bwd


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Polls SM clock / throttle reasons through NVML while the benchmark runs."""

    def __init__(self, index):
        self.samples = []
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), clk, reasons))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.ok:
            self.t.join(1.0)

    def summary(self, t0, t1):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed_region"
        if len(inside) < 3:  # region shorter than a few polls: fall back to everything sampled under load
            inside = self.samples
            window = "whole_run"
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        seen = set()
        for _, _, r in inside:
            for bit, nm in names.items():
                if r & bit:
                    seen.add(nm)
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": float(self.sm_max),
                "reasons": sorted(seen), "samples": len(inside), "window": window}


def make_inputs(B, rank):
    from neural_renderer_b200 import synthetic
    w = WORKLOAD
    faces = synthetic.sphere_faces(B, w["num_faces"], seed=1234 + 1000 * rank)
    tex = synthetic.random_textures(B, w["num_faces"], w["texture_size"], seed=4321 + rank)
    gen = torch.Generator().manual_seed(99 + rank)
    grad = torch.randn((B, 3, w["image_size"], w["image_size"]), generator=gen)
    return torch.from_numpy(faces), torch.from_numpy(tex), grad


# ------------------------------------------------------------------------------------------------ step functions
def ours_step(faces, tex, grad):
    import neural_renderer_b200 as nr
    w = WORKLOAD
    faces.grad = None
    tex.grad = None
    img = nr.rasterize(faces, tex, w["image_size"], w["anti_aliasing"], w["near"], w["far"], w["eps"], w["background"])
    img.backward(grad)  # upstream gradient dL/dI = grad, i.e. L = sum(I * grad)
    with torch.no_grad():
        loss = (img * grad).sum()
    return loss


class InputPipeline:
    """Double-buffered host -> device staging for the end-to-end measurement: the copy of step i+1's inputs (pinned
    host memory, its own stream) overlaps the kernels of step i.  Every step still pays for the copy of one full set
    of inputs inside the timed region (K steps issue K copies; the device-wide synchronize that closes the region
    waits for the last one)."""

    def __init__(self, dev, host_tensors):
        self.host = [t.pin_memory() for t in host_tensors]
        self.bufs = [[torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host_tensors] for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]  # copy into the slot has finished
        self.free = [torch.cuda.Event() for _ in range(2)]   # the step that used the slot has finished
        self.i = 0
        self._issue(0)

    def _issue(self, slot):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            for d, h in zip(self.bufs[slot], self.host):
                d.copy_(h, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def acquire(self):
        slot = self.i & 1
        torch.cuda.current_stream().wait_event(self.ready[slot])
        self._issue(slot ^ 1)  # start moving the next step's inputs
        return slot, self.bufs[slot]

    def release(self, slot):
        self.free[slot].record(torch.cuda.current_stream())
        self.i += 1


def ref_gpu_step(faces, tex, grad):
    import refhost
    w = WORKLOAD
    res = refhost.rasterize_rgbad(faces, tex, w["image_size"], w["anti_aliasing"], w["near"], w["far"], w["eps"],
                                  w["background"], True, False, False)
    gf, gt = res.backward(grad, None, None)  # same upstream gradient as the other arm
    loss = (res["rgb"] * grad).sum()
    return loss, gf, gt


def oracle_cpu_step(faces_np, tex_np, grad_np):
    import nr_oracle as o
    w = WORKLOAD
    res = o.rasterize_rgbad(faces_np, tex_np, w["image_size"], w["anti_aliasing"], w["near"], w["far"], w["eps"],
                            w["background"], True, False, False)
    loss = float((res["rgb"] * grad_np).sum())
    gf, gt = res.backward(grad_np, None, None)
    return loss, gf, gt


def shared_mesh_workload(args, world, rank, local_rank):
    """BASELINE.json configs[4] shape: ONE shared mesh rendered from many viewpoints, viewpoints sharded over the
    ranks, vertex / texture gradients summed across ranks with NCCL (the only collective this path has)."""
    import neural_renderer_b200 as nr
    from neural_renderer_b200 import synthetic
    from neural_renderer_b200.distributed import allreduce_shared_grads
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    barrier = (lambda: dist.barrier(device_ids=[local_rank])) if distributed else (lambda: None)
    F, S, ts, V = args.shared_faces, args.shared_image, 2, args.views_per_gpu
    v_np, f_np = synthetic.sphere_mesh(F)
    vertices = torch.from_numpy((v_np * 0.55).astype(np.float32)).to(dev).requires_grad_(True)   # shared parameters
    textures = torch.rand((F, ts, ts, ts, 3), generator=torch.Generator().manual_seed(7)).to(dev).requires_grad_(True)
    faces_idx = torch.from_numpy(f_np).to(dev)
    lo, hi = shard_range(world * V, rank, world)
    az = torch.arange(lo, hi, dtype=torch.float32) * (360.0 / (world * V))
    eyes = nr.get_points_from_angles(torch.full_like(az, 2.732), torch.full_like(az, 30.0), az).to(dev)
    renderer = nr.Renderer()
    renderer.image_size, renderer.anti_aliasing, renderer.fill_back = S, False, False
    renderer.eye = eyes
    grad = torch.randn((V, 3, S, S), generator=torch.Generator().manual_seed(99 + rank)).to(dev)

    def step():
        vertices.grad = None
        textures.grad = None
        img = renderer.render(vertices[None].expand(V, -1, -1), faces_idx[None].expand(V, -1, -1),
                              textures[None].expand(V, -1, -1, -1, -1, -1))
        (img * grad).sum().backward()
        for w in allreduce_shared_grads([vertices, textures], async_op=True):
            w.wait()

    ms, t0, t1 = timed_loop(step, args.steps, args.warmup, barrier)
    if distributed:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * V * S * S * args.steps / (ms * 1e-3) / 1e6
    if rank == 0:
        print(json.dumps({
            "metric": "Mpixels/s fwd+bwd, shared mesh, viewpoint-sharded", "value": round(value, 2), "unit": "Mpixels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "ours",
            "config": {"workload": "shared mesh: Renderer.render fwd+bwd, %d faces, %dx%d, ts=%d, %d views/GPU, sum "
                                   "all-reduce of vertex (%.1f MB) and texture (%.1f MB) gradients"
                                   % (F, S, S, ts, V, vertices.numel() * 4 / 1e6, textures.numel() * 4 / 1e6),
                       "collective": "nccl all_reduce(sum) x2 per step" if distributed else "none (1 rank)"}}))
    if distributed:
        dist.destroy_process_group()


def timed_loop(step, steps, warmup, barrier):
    """W warm-up steps, then exactly K steps bracketed by barrier + synchronize, timed with CUDA events on the
    current stream."""
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    barrier()
    return e0.elapsed_time(e1), t0, t1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=64, help="batch items of the workload timed on the CPU oracle")
    ap.add_argument("--no-side-measurements", action="store_true", help="skip cpu_baseline / reference_gpu / kernels")
    ap.add_argument("--workload", default="headline", choices=["headline", "shared_mesh"])
    ap.add_argument("--shared-faces", type=int, default=1000000)
    ap.add_argument("--shared-image", type=int, default=1024)
    ap.add_argument("--views-per-gpu", type=int, default=8)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    # stdout carries exactly one JSON line: NCCL's own banner / debug output (NCCL_DEBUG) goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    w = WORKLOAD
    B, F, S, ts = w["batch_per_gpu"], w["num_faces"], w["image_size"], w["texture_size"]
    sys.path.insert(0, os.path.join(ROOT, "oracle"))

    if args.impl == "reference":
        return reference_arm(args, world, rank, local_rank)
    if args.workload == "shared_mesh":
        return shared_mesh_workload(args, world, rank, local_rank)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA GPU: this package has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        barrier = lambda: dist.barrier(device_ids=[local_rank])  # noqa: E731
    else:
        barrier = lambda: None  # noqa: E731

    from neural_renderer_b200 import _lib
    lib = _lib.load()  # fails loudly when libnr_b200.so is missing

    faces_h, tex_h, grad_h = make_inputs(B, rank)
    faces = faces_h.to(dev).requires_grad_(True)
    tex = tex_h.to(dev).requires_grad_(True)
    grad = grad_h.to(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- headline: device-resident inputs
    launches_per_step = [0]

    def step():
        ours_step(faces, tex, grad)

    ms, t0, t1 = timed_loop(step, args.steps, args.warmup, barrier)
    if distributed:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pixels = world * B * S * S
    value = pixels * args.steps / (ms * 1e-3) / 1e6
    clocks = sampler.summary(t0, t1)

    # count our kernel launches of one step through the library's own accounting
    img = None
    import neural_renderer_b200 as nr
    faces.grad = None
    tex.grad = None
    img = nr.rasterize(faces, tex, S, False, w["near"], w["far"], w["eps"], w["background"])
    n_fwd = lib.nr_b200_last_launch_count()
    (img * grad).sum().backward()
    n_bwd = lib.nr_b200_last_launch_count()
    launches_per_step[0] = n_fwd + n_bwd

    # ---- end to end: host (pinned) inputs in, loss + vertex gradients out, copies inside the timed region.
    #      Headline e2e: every step copies ITS OWN inputs, computes, reads back -- strictly in sequence.
    faces_p, tex_p = faces_h.pin_memory(), tex_h.pin_memory()
    gf_host = torch.empty_like(faces_h).pin_memory()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        f = faces_p.to(dev, non_blocking=True).requires_grad_(True)
        t = tex_p.to(dev, non_blocking=True).requires_grad_(True)
        loss = ours_step(f, t, grad)
        loss_host.copy_(loss.detach(), non_blocking=True)
        gf_host.copy_(f.grad, non_blocking=True)

    e2e_ms, _, _ = timed_loop(e2e_step, args.steps, args.warmup, barrier)
    if distributed:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = pixels * args.steps / (e2e_ms * 1e-3) / 1e6
    h2d = faces_h.numel() * 4 + tex_h.numel() * 4
    d2h = gf_host.numel() * 4 + 4

    #      Reported beside it: the same loop with double-buffered staging (the copy of step i+1's inputs overlaps the
    #      kernels of step i; still one full copy per step inside the timed region).
    pipe = InputPipeline(dev, [faces_h, tex_h])

    def e2e_pipelined_step():
        slot, (f_buf, t_buf) = pipe.acquire()
        f = f_buf.detach().requires_grad_(True)
        t = t_buf.detach().requires_grad_(True)
        loss = ours_step(f, t, grad)
        loss_host.copy_(loss.detach(), non_blocking=True)
        gf_host.copy_(f.grad, non_blocking=True)
        pipe.release(slot)

    pipe_ms, _, _ = timed_loop(e2e_pipelined_step, args.steps, args.warmup, barrier)
    if distributed:
        t = torch.tensor([pipe_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pipe_ms = float(t.item())
    pipe_value = pixels * args.steps / (pipe_ms * 1e-3) / 1e6

    out = {
        "metric": "Mpixels/s fwd+bwd @ 256x256, 5k faces, batch 64", "value": round(value, 2), "unit": "Mpixels/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": "ours",
        "config": {"workload": "headline: rasterize() fwd+bwd RGB, B=%d/GPU x F=%d faces, %dx%d, ts=%d, "
                               "anti_aliasing off (BASELINE.json metric; configs[3] batch-sharded shape per GPU)"
                               % (B, F, S, S, ts),
                   "global_batch": world * B, "num_faces": F, "image_size": S, "texture_size": ts,
                   "parallelism": "batch-sharded x%d (no data-path collective)" % world,
                   "l2": "no explicit flush: per-step working set (textures 245.8 MB + grad_textures 245.8 MB + "
                         "maps 134 MB) exceeds the 126 MB L2",
                   "upstream_grad": "dense N(0,1), seed 99"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": "Mpixels/s", "ms_per_step": round(e2e_ms / args.steps, 4),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "per step, in sequence: pinned host faces+textures -> device, rasterize fwd+bwd, loss + grad_faces -> host",
                "pipelined": {"value": round(pipe_value, 2), "unit": "Mpixels/s", "ms_per_step": round(pipe_ms / args.steps, 4),
                              "what": "same work with double-buffered staging: the copy of the next step's inputs overlaps "
                                      "this step's kernels (one full copy per step inside the timed region)"}},
        "gpu_launches": launches_per_step[0] * args.steps,
        "gpu_launches_per_step": launches_per_step[0],
    }

    # ---- side measurements (rank 0, outside the headline region)
    if rank == 0 and not args.no_side_measurements:
        peak, peak_src = measured_peaks()
        fwd_bytes, bwd_bytes = algorithmic_bytes(B, F, S, ts)
        lib.nr_b200_set_profiling(1)
        _lib.read_profile()
        nprof = max(5, min(args.steps, 20))
        for _ in range(nprof):
            ours_step(faces, tex, grad)
        torch.cuda.synchronize()
        prof = _lib.read_profile()
        lib.nr_b200_set_profiling(0)
        per = {}
        for name, v in prof:
            per.setdefault(name, []).append(v)
        # a name can appear twice per step (k_face_bbox runs in both passes): report per-step totals
        kern = {k: round(sum(v) / nprof, 5) for k, v in per.items()}
        out["kernels_ms_per_step"] = kern
        fwd_ms = kern.get("k_raster_tile", 0.0)
        dom = max(kern, key=lambda k: kern[k]) if kern else None

        traffic = {}
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this shape (profiles/)
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f)["bytes_per_launch"]
        except Exception:
            pass

        def roof(bytes_, ms_, kernel=None):
            ach = bytes_ / (ms_ * 1e-3) / 1e9 if ms_ > 0 else 0.0
            return {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic.get(kernel), "peak_source": peak_src,
                    "algorithmic_bytes": bytes_, "kernel_ms": round(ms_, 5)}

        if dom:
            dom_bytes = fwd_bytes if dom == "k_raster_tile" else bwd_bytes
            out["roofline"] = dict(roof(dom_bytes, kern[dom], dom), kernel=dom,
                                   note="algorithmic bytes of the pass the kernel belongs to (SURVEY.md 8(d)) / "
                                        "kernel duration; an ALU/atomic-bound kernel reads far below the HBM roof")
        out["roofline_fwd"] = dict(roof(fwd_bytes, fwd_ms + kern.get("k_face_bbox", 0.0) / 2, "k_raster_tile"), kernel="k_raster_tile",
                                   note="forward rasterize = k_face_bbox + k_raster_tile, 391.5 MB algorithmic")

        # reference's own kernels on this GPU (the reported baseline of BASELINE.md section 2); N = 1 only
        try:
            if world > 1:
                raise RuntimeError("reported at N=1 only")
            import refhost
            if refhost.available(S, F, ts, w["near"], w["far"], w["eps"], 1, 0, 0):
                fr, tr = faces.detach(), tex.detach()
                rms, _, _ = timed_loop(lambda: ref_gpu_step(fr, tr, grad), max(3, args.steps // 4), 3, lambda: None)
                rsteps = max(3, args.steps // 4)
                out["reference_gpu"] = {"value": round(B * S * S * rsteps / (rms * 1e-3) / 1e6, 2),
                                        "unit": "Mpixels/s", "ms_per_step": round(rms / rsteps, 3), "steps": rsteps,
                                        "what": "reference CuPy kernel strings re-hosted (oracle/_ref), same inputs, 1 GPU"}
            else:
                out["reference_gpu"] = {"unavailable": "oracle/_ref binaries for the headline shape not built"}
        except Exception as e:  # pragma: no cover
            out["reference_gpu"] = {"unavailable": repr(e)[:200]}

        # CPU oracle (port) on a bounded sample of the same workload; N = 1 only (torchrun pins OMP threads to 1)
        try:
            if world > 1:
                raise RuntimeError("reported at N=1 only")
            import nr_oracle as o
            nb = max(1, min(B, args.cpu_sample))
            fn, tn, gn = faces_h[:nb].numpy(), tex_h[:nb].numpy(), grad_h[:nb].numpy()
            tc0 = time.perf_counter()
            oracle_cpu_step(fn, tn, gn)
            tc = time.perf_counter() - tc0
            out["cpu_baseline"] = {"value": round(nb * S * S / tc / 1e6, 4), "unit": "Mpixels/s",
                                   "cores": o.num_threads(), "kind": "port",
                                   "sample": "%d of the %d batch items, fwd+bwd, %.1f s" % (nb, B, tc)}
        except Exception as e:  # pragma: no cover
            out["cpu_baseline"] = {"unavailable": repr(e)[:200]}

    sampler.stop()
    if rank == 0:
        print(json.dumps(out))
    if distributed:
        dist.destroy_process_group()


def reference_arm(args, world, rank, local_rank):
    """The reference's own implementation of the path, same workload / metric (rank 0 only)."""
    if rank != 0:
        return
    w = WORKLOAD
    B, F, S, ts = w["batch_per_gpu"], w["num_faces"], w["image_size"], w["texture_size"]
    faces_h, tex_h, grad_h = make_inputs(B, 0)
    base = {"metric": "Mpixels/s fwd+bwd @ 256x256, 5k faces, batch 64", "unit": "Mpixels/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference"}
    use_gpu = False
    if torch.cuda.is_available():
        import refhost
        use_gpu = refhost.available(S, F, ts, w["near"], w["far"], w["eps"], 1, 0, 0)
    if use_gpu:
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        faces, tex, grad = faces_h.to(dev), tex_h.to(dev), grad_h.to(dev)
        sampler = ClockSampler(local_rank)
        sampler.start()
        ms, t0, t1 = timed_loop(lambda: ref_gpu_step(faces, tex, grad), args.steps, args.warmup, lambda: None)
        value = B * S * S * args.steps / (ms * 1e-3) / 1e6
        faces_p, tex_p = faces_h.pin_memory(), tex_h.pin_memory()
        gf_host = torch.empty_like(faces_h).pin_memory()
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()

        def e2e_step():
            f = faces_p.to(dev, non_blocking=True)
            t = tex_p.to(dev, non_blocking=True)
            loss, gf, _ = ref_gpu_step(f, t, grad)
            loss_host.copy_(loss, non_blocking=True)
            gf_host.copy_(gf, non_blocking=True)

        e2e_ms, _, _ = timed_loop(e2e_step, args.steps, args.warmup, lambda: None)
        e2e_value = B * S * S * args.steps / (e2e_ms * 1e-3) / 1e6
        pipe = InputPipeline(dev, [faces_h, tex_h])  # same double-buffered staging as the other arm

        def e2e_pipelined_step():
            slot, (f, t) = pipe.acquire()
            loss, gf, _ = ref_gpu_step(f, t, grad)
            loss_host.copy_(loss, non_blocking=True)
            gf_host.copy_(gf, non_blocking=True)
            pipe.release(slot)

        pipe_ms, _, _ = timed_loop(e2e_pipelined_step, args.steps, args.warmup, lambda: None)
        pipe_value = B * S * S * args.steps / (pipe_ms * 1e-3) / 1e6
        clocks = sampler.summary(t0, t1)
        sampler.stop()
        base.update({
            "value": round(value, 2), "ms_per_step": round(ms / args.steps, 4), "clocks": clocks,
            "config": {"workload": "headline: reference kernels (K1,K2,K4,K5,K6 of rasterize.py, unmodified strings "
                                   "re-hosted without CuPy) fwd+bwd RGB, B=%d x F=%d, %dx%d, ts=%d" % (B, F, S, S, ts),
                       "device": "cuda (the reference ships no CPU implementation: rasterize.py:893-897)",
                       "global_batch": B, "num_faces": F, "image_size": S, "texture_size": ts},
            "cpu_baseline": {"value": round(value, 2), "unit": "Mpixels/s", "cores": 0, "kind": "reference",
                             "sample": "full workload on the GPU: the reference has no CPU path, its own CUDA kernels "
                                       "are the baseline (oracle/_ref)"},
            "e2e": {"value": round(e2e_value, 2), "unit": "Mpixels/s", "ms_per_step": round(e2e_ms / args.steps, 4),
                    "h2d_bytes_per_step": faces_h.numel() * 4 + tex_h.numel() * 4,
                    "d2h_bytes_per_step": gf_host.numel() * 4 + 4,
                    "pipelined": {"value": round(pipe_value, 2), "unit": "Mpixels/s",
                                  "ms_per_step": round(pipe_ms / args.steps, 4)}},
            "gpu_launches": 0,
        })
    else:
        import nr_oracle as o
        nb = max(1, min(B, args.cpu_sample))
        fn, tn, gn = faces_h[:nb].numpy(), tex_h[:nb].numpy(), grad_h[:nb].numpy()
        for _ in range(min(args.warmup, 1)):
            oracle_cpu_step(fn[:1], tn[:1], gn[:1])
        t0 = time.perf_counter()
        steps = max(1, min(args.steps, 3))
        for _ in range(steps):
            oracle_cpu_step(fn, tn, gn)
        dt = (time.perf_counter() - t0) / steps
        value = nb * S * S / dt / 1e6
        base.update({
            "value": round(value, 4), "steps": steps, "ms_per_step": round(dt * 1e3, 2),
            "config": {"workload": "headline sample on the CPU oracle port (oracle/_ref binaries absent)",
                       "global_batch": nb, "num_faces": F, "image_size": S, "texture_size": ts},
            "cpu_baseline": {"value": round(value, 4), "unit": "Mpixels/s", "cores": o.num_threads(), "kind": "port",
                             "sample": "%d of %d batch items per step" % (nb, B)},
            "e2e": {"value": round(value, 4), "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        })
    print(json.dumps(base))


if __name__ == "__main__":
    main()
