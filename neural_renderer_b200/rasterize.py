"""Host-side mirror of the reference's rasterizer interface (neural_renderer/rasterize.py) on top of the C ABI.

Same names, argument order, defaults and error conditions as the reference:

  rasterize_rgbad        rasterize.py:900-977
  rasterize              rasterize.py:980-1008
  rasterize_silhouettes  rasterize.py:1011-1034
  rasterize_depth        rasterize.py:1037-1060
  Rasterize              rasterize.py:19-897   (function object; returns un-flipped NHWC maps like forward_gpu)
  use_unsafe_rasterizer  rasterize.py:1063-1065

Tensors are CUDA `torch.Tensor`s instead of chainer Variables / cupy arrays; PyTorch only provides device memory,
the current stream and autograd bookkeeping -- all arithmetic happens in libnr_b200.so (hand-written sm_100a CUDA).
There is no CPU path (the reference raises NotImplementedError for CPU arrays as well, rasterize.py:893-897).
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib

DEFAULT_IMAGE_SIZE = 256
DEFAULT_ANTI_ALIASING = True
DEFAULT_NEAR = 0.1
DEFAULT_FAR = 100
DEFAULT_EPS = 1e-4
DEFAULT_BACKGROUND_COLOR = (0, 0, 0)
USE_UNSAFE_IMPLEMENTATION = False

# rasterize.py:389 fetches the vertex depths for texture sampling from batch item 0.  Reference-exact by default;
# NEURAL_RENDERER_B200_FIX_TEXTURE_DEPTH=1 (or set_reference_exact(False)) samples with each item's own depths.
_REFERENCE_EXACT = not int(os.environ.get("NEURAL_RENDERER_B200_FIX_TEXTURE_DEPTH", "0"))
# NR_FWD_STAGE_TEXTURES (texture cubes staged in shared memory with cp.async.bulk): same pixels, measured slower than the
# direct gather on B200 -- kept selectable for measurements and tests, off by default.
_STAGE_TEXTURES = False


def set_stage_textures(flag):
    global _STAGE_TEXTURES
    _STAGE_TEXTURES = bool(flag)


def set_reference_exact(flag):
    global _REFERENCE_EXACT
    _REFERENCE_EXACT = bool(flag)


def use_unsafe_rasterizer(flag):
    """Accepted for interface compatibility (rasterize.py:1063).  The reference's 'unsafe' scanline/spin-lock
    variant is an alternative implementation of the same maps with arrival-order tie breaks; this package has a
    single deterministic forward path, so the flag changes nothing."""
    global USE_UNSAFE_IMPLEMENTATION
    USE_UNSAFE_IMPLEMENTATION = bool(flag)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check_inputs(faces, textures, return_rgb, face_light=None, textures_fill_back=False, vertices=None):
    # rasterize.py:66-90 (chainer type_check) -> TypeError / ValueError with the same conditions
    if not isinstance(faces, torch.Tensor):
        raise TypeError("faces must be a torch.Tensor")
    if vertices is not None:
        # indexed geometry (vertices_to_faces.py:10-14 asserts): vertices [B,Nv,3] float, faces [B,F,3] / [1,F,3] / [F,3] int
        if not isinstance(vertices, torch.Tensor) or not vertices.is_floating_point():
            raise TypeError("vertices must be a floating point torch.Tensor")
        if vertices.dim() != 3 or vertices.shape[2] != 3:
            raise ValueError("vertices must have shape [batch size, num vertices, 3], got %s" % (tuple(vertices.shape),))
        if faces.is_floating_point():
            raise TypeError("with `vertices`, faces must hold integer vertex indices")
        if not ((faces.dim() == 2 and faces.shape[1] == 3)
                or (faces.dim() == 3 and faces.shape[2] == 3 and faces.shape[0] in (1, vertices.shape[0]))):
            raise ValueError("with `vertices`, faces must have shape [batch size, num faces, 3] or [num faces, 3], got %s"
                             % (tuple(faces.shape),))
        if not vertices.is_cuda or not faces.is_cuda:
            raise NotImplementedError("neural_renderer_b200 has no CPU implementation (inputs must be CUDA tensors)")
        batch_size, num_faces = vertices.shape[0], faces.shape[-2]
    else:
        if not faces.is_floating_point():
            raise TypeError("faces must be floating point")
        if faces.dim() != 4 or faces.shape[2] != 3 or faces.shape[3] != 3:
            raise ValueError("faces must have shape [batch size, num faces, 3, 3], got %s" % (tuple(faces.shape),))
        batch_size, num_faces = faces.shape[0], faces.shape[1]
    if return_rgb:
        if not isinstance(textures, torch.Tensor):
            raise TypeError("textures are required to draw RGB")
        if not textures.is_floating_point():
            raise TypeError("textures must be floating point")
        num_cubes = num_faces // 2 if textures_fill_back else num_faces
        if textures_fill_back and num_faces % 2:
            raise ValueError("textures_fill_back needs an even number of faces (front faces, then their reversed copies)")
        # batch size 1 with a larger geometry batch = ONE set of cubes shared by every item (a mesh seen from B viewpoints)
        if (textures.dim() != 6 or textures.shape[2] < 2 or textures.shape[2] != textures.shape[3]
                or textures.shape[3] != textures.shape[4] or textures.shape[5] != 3
                or textures.shape[0] not in (1, batch_size) or textures.shape[1] != num_cubes):
            raise ValueError("textures must have shape [batch size, num faces, ts, ts, ts, 3] with ts >= 2 and match "
                             "faces, got %s" % (tuple(textures.shape),))
    if return_rgb and face_light is not None:
        if not isinstance(face_light, torch.Tensor) or tuple(face_light.shape) != (batch_size, num_faces, 3):
            raise ValueError("face_light must have shape [batch size, num faces, 3]")
        if not face_light.is_cuda:
            raise NotImplementedError("neural_renderer_b200 has no CPU implementation (inputs must be CUDA tensors)")
    if not faces.is_cuda or (return_rgb and not textures.is_cuda):
        raise NotImplementedError("neural_renderer_b200 has no CPU implementation (inputs must be CUDA tensors)")


# Optional hook between the two halves of the backward pass (neural_renderer_b200.distributed.overlap_texture_allreduce):
# called as hook(grad_textures) right after the texture-gradient kernels are enqueued and BEFORE the edge scan is;
# returns an object whose .wait() is called once the whole pass is enqueued.
_TEXTURE_GRAD_HOOK = None


def set_texture_grad_hook(hook):
    global _TEXTURE_GRAD_HOOK
    prev = _TEXTURE_GRAD_HOOK
    _TEXTURE_GRAD_HOOK = hook
    return prev


class _Config:
    __slots__ = ("S", "aa", "near", "far", "eps", "bg", "bg_batch", "flags", "reference_exact")


def _make_config(image_size, anti_aliasing, near, far, eps, background_color, return_rgb, return_alpha, return_depth,
                 device, batch_size, reference_exact=None):
    if not any((return_rgb, return_alpha, return_depth)):
        raise Exception("nothing to draw")  # rasterize.py:25-27 raises a bare Exception
    cfg = _Config()
    cfg.aa = bool(anti_aliasing)
    cfg.S = int(image_size) * 2 if cfg.aa else int(image_size)
    cfg.near, cfg.far, cfg.eps = float(near), float(far), float(eps)
    flags = 0
    if return_rgb:
        flags |= _lib.NR_RETURN_RGB
    if return_alpha:
        flags |= _lib.NR_RETURN_ALPHA
    if return_depth:
        flags |= _lib.NR_RETURN_DEPTH
    if cfg.aa:
        flags |= _lib.NR_ANTI_ALIASING
    cfg.reference_exact = _REFERENCE_EXACT if reference_exact is None else bool(reference_exact)
    if cfg.reference_exact:
        flags |= _lib.NR_TEX_Z_BATCH0
    cfg.bg = (0.0, 0.0, 0.0)
    cfg.bg_batch = None
    if return_rgb:
        bg = background_color
        if isinstance(bg, torch.Tensor):
            bg = bg.detach().to(dtype=torch.float32)
        else:
            bg = torch.as_tensor(bg, dtype=torch.float32)
        if bg.dim() == 1 and bg.numel() == 3:
            cfg.bg = tuple(float(v) for v in bg.tolist())
        elif bg.dim() == 2 and bg.shape[1] == 3 and bg.shape[0] == batch_size:  # rasterize.py:464-465
            cfg.bg_batch = bg.to(device).contiguous()
            flags |= _lib.NR_BG_PER_BATCH
        else:
            raise ValueError("background_color must have shape (3,) or (batch size, 3)")
    cfg.flags = flags
    return cfg


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _RasterizeFunction(torch.autograd.Function):
    """autograd node of the hot path: forward = nr_b200_forward, backward = nr_b200_backward.

    `geom` is faces [B,F,3,3], or -- indexed geometry, `indices` given -- vertices [B,Nv,3] (the vertices_to_faces
    gather and its scatter-add backward happen inside the kernels).  Outputs are the API images (planar, image
    orientation, pooled when anti-aliasing) plus the raster-resolution maps (returned non-differentiable so tests /
    the `Rasterize` object can look at them)."""

    @staticmethod
    def forward(ctx, geom, textures, face_light, cfg, indices):
        lib = _lib.load()
        dev = geom.device
        geom_c = geom.detach().contiguous()
        tex_c = textures.detach().contiguous() if textures is not None else None
        light_c = face_light.detach().to(torch.float32).contiguous() if face_light is not None else None
        flags = cfg.flags
        if indices is not None:
            B, Nv = geom_c.shape[:2]
            F = indices.shape[-2]
            flags |= _lib.NR_FACES_INDEXED
            if indices.dim() == 2 or (indices.shape[0] == 1 and B > 1):
                flags |= _lib.NR_INDICES_SHARED
        else:
            B, F = geom_c.shape[:2]
            Nv = 0
        if tex_c is not None and tex_c.shape[0] == 1 and B > 1:
            flags |= _lib.NR_TEX_SHARED
        S = cfg.S
        ts = int(tex_c.shape[2]) if tex_c is not None else 0
        want_rgb = bool(flags & _lib.NR_RETURN_RGB)
        want_alpha = bool(flags & _lib.NR_RETURN_ALPHA)
        want_depth = bool(flags & _lib.NR_RETURN_DEPTH)
        with torch.cuda.device(dev):
            fim = torch.empty((B, S, S), dtype=torch.int32, device=dev)
            wmap = torch.empty((B, 3, S, S), dtype=torch.float32, device=dev)
            dmap = torch.empty((B, S, S), dtype=torch.float32, device=dev)
            rgb_map = torch.empty((B, 3, S, S), dtype=torch.float32, device=dev) if want_rgb else None
            alpha_map = torch.empty((B, S, S), dtype=torch.float32, device=dev) if want_alpha else None
            out_rgb = out_alpha = out_depth = None
            if cfg.aa:
                H = S // 2
                if want_rgb:
                    out_rgb = torch.empty((B, 3, H, H), dtype=torch.float32, device=dev)
                if want_alpha:
                    out_alpha = torch.empty((B, H, H), dtype=torch.float32, device=dev)
                if want_depth:
                    out_depth = torch.empty((B, H, H), dtype=torch.float32, device=dev)
            ws_bytes = lib.nr_b200_forward_workspace_bytes(B, F, S, ts, flags)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            a = _lib.ForwardArgs()
            a.struct_size = ctypes.sizeof(_lib.ForwardArgs)
            a.flags = flags
            a.batch_size, a.num_faces, a.raster_size, a.texture_size = B, F, S, ts
            a.near_, a.far_, a.eps = cfg.near, cfg.far, cfg.eps
            a.background[0], a.background[1], a.background[2] = cfg.bg
            if indices is not None:
                a.vertices, a.face_indices, a.num_vertices = _ptr(geom_c), _ptr(indices), Nv
            else:
                a.faces = _ptr(geom_c)
            a.textures, a.background_batch = _ptr(tex_c), _ptr(cfg.bg_batch)
            a.face_index_map, a.weight_map, a.depth_map = _ptr(fim), _ptr(wmap), _ptr(dmap)
            a.rgb_map, a.alpha_map = _ptr(rgb_map), _ptr(alpha_map)
            a.out_rgb, a.out_alpha, a.out_depth = _ptr(out_rgb), _ptr(out_alpha), _ptr(out_depth)
            a.workspace, a.workspace_bytes = _ptr(ws), ws_bytes
            a.face_light = _ptr(light_c)
            _lib.check(lib.nr_b200_forward(ctypes.byref(a), _stream_ptr(dev)))
        ctx.cfg = cfg
        ctx.flags = flags
        ctx.ts = ts
        ctx.F = F
        ctx.tex_shape = tuple(textures.shape) if textures is not None else None
        # the unlit textures are only needed again for d loss / d face_light
        need_light_grad = light_c is not None and ctx.needs_input_grad[2]
        ctx.save_for_backward(geom_c, fim, wmap, dmap, rgb_map, light_c, tex_c if need_light_grad else None, indices)
        if cfg.aa:
            rgb_o, alpha_o, depth_o = out_rgb, out_alpha, out_depth
        else:
            rgb_o, alpha_o, depth_o = rgb_map, alpha_map, (dmap if want_depth else None)
        ctx.mark_non_differentiable(fim, wmap)
        return rgb_o, alpha_o, depth_o, fim, wmap

    @staticmethod
    def backward(ctx, g_rgb, g_alpha, g_depth, _g_fim, _g_wmap):
        lib = _lib.load()
        cfg = ctx.cfg
        flags = ctx.flags
        geom_c, fim, wmap, dmap, rgb_map, light_c, tex_c, indices = ctx.saved_tensors
        dev = geom_c.device
        B, F = geom_c.shape[0], ctx.F
        want_rgb = bool(flags & _lib.NR_RETURN_RGB)

        def prep(g, wanted):
            if g is None or not wanted:
                return None
            return g.detach().to(torch.float32).contiguous()

        g_rgb = prep(g_rgb, want_rgb)
        g_alpha = prep(g_alpha, bool(flags & _lib.NR_RETURN_ALPHA))
        g_depth = prep(g_depth, bool(flags & _lib.NR_RETURN_DEPTH))
        with torch.cuda.device(dev):
            grad_geom = torch.empty_like(geom_c)  # grad_faces [B,F,3,3], or grad_vertices [B,Nv,3] when indexed
            grad_textures = torch.empty(ctx.tex_shape, dtype=torch.float32, device=dev) if want_rgb else None
            grad_light = torch.empty_like(light_c) if (want_rgb and tex_c is not None) else None
            ws_bytes = lib.nr_b200_backward_workspace_bytes(B, F, cfg.S, ctx.ts, flags)
            ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
            a = _lib.BackwardArgs()
            a.struct_size = ctypes.sizeof(_lib.BackwardArgs)
            a.batch_size, a.num_faces, a.raster_size, a.texture_size = B, F, cfg.S, ctx.ts
            a.eps = cfg.eps
            if indices is not None:
                a.vertices, a.face_indices, a.num_vertices = _ptr(geom_c), _ptr(indices), geom_c.shape[1]
                a.grad_vertices = _ptr(grad_geom)
            else:
                a.faces, a.grad_faces = _ptr(geom_c), _ptr(grad_geom)
            a.textures = _ptr(tex_c)
            a.face_light, a.grad_face_light = _ptr(light_c), _ptr(grad_light)
            a.face_index_map, a.weight_map, a.depth_map, a.rgb_map = _ptr(fim), _ptr(wmap), _ptr(dmap), _ptr(rgb_map)
            a.grad_rgb, a.grad_alpha, a.grad_depth = _ptr(g_rgb), _ptr(g_alpha), _ptr(g_depth)
            a.grad_textures = _ptr(grad_textures)
            a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
            hook = _TEXTURE_GRAD_HOOK if (want_rgb and g_rgb is not None) else None
            if hook is None:
                a.flags = flags
                _lib.check(lib.nr_b200_backward(ctypes.byref(a), _stream_ptr(dev)))
            else:
                # two halves: the texture gradient is complete (and may start its all-reduce on another stream)
                # before the edge scan is even enqueued
                a.flags = flags | _lib.NR_BWD_PART_TEXTURES
                _lib.check(lib.nr_b200_backward(ctypes.byref(a), _stream_ptr(dev)))
                pending = hook(grad_textures)
                a.flags = flags | _lib.NR_BWD_PART_FACES
                _lib.check(lib.nr_b200_backward(ctypes.byref(a), _stream_ptr(dev)))
                if pending is not None:
                    pending.wait()
        return grad_geom, grad_textures, grad_light, None, None


def _run(faces, textures, image_size, anti_aliasing, near, far, eps, background_color, return_rgb, return_alpha,
         return_depth, face_light=None, textures_fill_back=False, vertices=None, reference_exact=None):
    _check_inputs(faces, textures, return_rgb, face_light, textures_fill_back, vertices)
    indices = None
    if vertices is not None:
        geom = vertices if vertices.dtype == torch.float32 else vertices.float()
        indices = faces
        if indices.dim() == 3 and indices.shape[0] > 1 and indices.stride(0) == 0:
            indices = indices[:1]  # an expanded [F,3] index set (Mesh.get_batch): keep it shared, do not materialise
        indices = indices.to(torch.int32).contiguous()
    else:
        geom = faces if faces.dtype == torch.float32 else faces.float()
    batch_size = geom.shape[0]
    if return_rgb:
        if textures.dtype != torch.float32:
            textures = textures.float()
        if textures.shape[0] == batch_size > 1 and textures.stride(0) == 0:
            textures = textures[:1]  # an expanded shared texture set: sample it in place (NR_TEX_SHARED)
    cfg = _make_config(image_size, anti_aliasing, near, far, eps, background_color, return_rgb, return_alpha,
                       return_depth, geom.device, batch_size, reference_exact)
    if return_rgb and textures_fill_back:
        cfg.flags |= _lib.NR_TEX_FILL_BACK
    if return_rgb and _STAGE_TEXTURES:
        cfg.flags |= _lib.NR_FWD_STAGE_TEXTURES
    return _RasterizeFunction.apply(geom, textures if return_rgb else None, face_light if return_rgb else None, cfg,
                                    indices)


def rasterize_rgbad(
        faces,
        textures=None,
        image_size=DEFAULT_IMAGE_SIZE,
        anti_aliasing=DEFAULT_ANTI_ALIASING,
        near=DEFAULT_NEAR,
        far=DEFAULT_FAR,
        eps=DEFAULT_EPS,
        background_color=DEFAULT_BACKGROUND_COLOR,
        return_rgb=True,
        return_alpha=True,
        return_depth=True,
        *,
        face_light=None,
        textures_fill_back=False,
        vertices=None,
        reference_exact=None,
):
    """Generate RGB, alpha channel, and depth images from faces and textures (for RGB).  rasterize.py:900-977.

    Returns {'rgb': [B,3,H,W], 'alpha': [B,H,W], 'depth': [B,H,W]} (None for the ones not requested).

    Keyword-only extensions (not in the reference; `Renderer.render` uses them so that neither the lit nor the
    fill_back-doubled texture tensor is ever materialised):
      face_light [B,F,3]      per-face RGB factor of `lighting` applied at sample time (== sampling textures * light)
      textures_fill_back      faces [F/2, F) are the reversed copies of [0, F/2) and `textures` holds only the F/2
                              original cubes (the copies read them with reversed axes, renderer.py:80)
      vertices [B,Nv,3]       indexed geometry: `faces` then holds integer vertex indices ([B,F,3], or [F,3] / [1,F,3]
                              shared by the batch) and vertices_to_faces (vertices_to_faces.py:16-21) plus its
                              scatter-add backward run inside the rasterizer: no [B,F,3,3] tensor exists and the
                              gradient arrives in `vertices.grad`
      reference_exact         True / False overrides the module default (`set_reference_exact`) for this call: the
                              reference's texture sampler reads the vertex depths of batch item 0 for EVERY item
                              (rasterize.py:389).  True reproduces that bit for bit; False samples every item with its own
                              depths -- what one wants for batches of different meshes or cameras (the images of items
                              b > 0 and grad_textures differ, item 0 and all silhouettes / depths do not)
    `textures` with batch size 1 (or an expanded stride-0 batch) while the geometry batch is larger = one texture set
    shared by every item (a mesh seen from B viewpoints, mesh.py:29-34); its gradient is the sum over the items."""
    rgb, alpha, depth, _, _ = _run(faces, textures, image_size, anti_aliasing, near, far, eps, background_color,
                                   return_rgb, return_alpha, return_depth, face_light, textures_fill_back, vertices,
                                   reference_exact)
    return {
        'rgb': rgb if return_rgb else None,
        'alpha': alpha if return_alpha else None,
        'depth': depth if return_depth else None,
    }


def rasterize(
        faces,
        textures,
        image_size=DEFAULT_IMAGE_SIZE,
        anti_aliasing=DEFAULT_ANTI_ALIASING,
        near=DEFAULT_NEAR,
        far=DEFAULT_FAR,
        eps=DEFAULT_EPS,
        background_color=DEFAULT_BACKGROUND_COLOR,
        *,
        face_light=None,
        textures_fill_back=False,
        vertices=None,
        reference_exact=None,
):
    """RGB images [B,3,H,W] from faces and textures.  rasterize.py:980-1008 (keyword-only extras: rasterize_rgbad)."""
    return rasterize_rgbad(
        faces, textures, image_size, anti_aliasing, near, far, eps, background_color, True, False, False,
        face_light=face_light, textures_fill_back=textures_fill_back, vertices=vertices,
        reference_exact=reference_exact)['rgb']


def rasterize_silhouettes(
        faces,
        image_size=DEFAULT_IMAGE_SIZE,
        anti_aliasing=DEFAULT_ANTI_ALIASING,
        near=DEFAULT_NEAR,
        far=DEFAULT_FAR,
        eps=DEFAULT_EPS,
        *,
        vertices=None,
):
    """Alpha channels [B,H,W] from faces.  rasterize.py:1011-1034 (keyword-only `vertices`: rasterize_rgbad)."""
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None, False, True, False,
                           vertices=vertices)['alpha']


def rasterize_depth(
        faces,
        image_size=DEFAULT_IMAGE_SIZE,
        anti_aliasing=DEFAULT_ANTI_ALIASING,
        near=DEFAULT_NEAR,
        far=DEFAULT_FAR,
        eps=DEFAULT_EPS,
        *,
        vertices=None,
):
    """Depth images [B,H,W] from faces.  rasterize.py:1037-1060 (keyword-only `vertices`: rasterize_rgbad)."""
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None, False, False, True,
                           vertices=vertices)['depth']


class Rasterize(object):
    """The reference's function object (rasterize.py:19-64): `Rasterize(image_size, near, far, eps,
    background_color, return_rgb, return_alpha, return_depth)(faces[, textures]) -> (rgb, alpha, depth)` with the
    reference's internal conventions (NHWC rgb, rows NOT flipped, None for outputs not requested).  After the call
    `face_index_map` / `weight_map` hold the raster maps in those same conventions (as the reference instance does)."""

    def __init__(self, image_size, near, far, eps, background_color, return_rgb=False, return_alpha=False,
                 return_depth=False):
        if not any((return_rgb, return_alpha, return_depth)):
            raise Exception("nothing to draw")
        self.image_size = image_size
        self.near = near
        self.far = far
        self.eps = eps
        self.background_color = background_color
        self.return_rgb = return_rgb
        self.return_alpha = return_alpha
        self.return_depth = return_depth
        self.face_index_map = None
        self.weight_map = None

    def __call__(self, faces, textures=None):
        rgb, alpha, depth, fim, wmap = _run(faces, textures, self.image_size, False, self.near, self.far, self.eps,
                                            self.background_color, self.return_rgb, self.return_alpha,
                                            self.return_depth)
        self.face_index_map = fim.flip(1)
        self.weight_map = wmap.permute(0, 2, 3, 1).flip(1)
        rgb = rgb.permute(0, 2, 3, 1).flip(1) if self.return_rgb else None
        alpha = alpha.flip(1) if self.return_alpha else None
        depth = depth.flip(1) if self.return_depth else None
        return rgb, alpha, depth
