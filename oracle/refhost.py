"""GPU re-host of the reference's own CuPy path, without CuPy/Chainer (TEST INFRASTRUCTURE / timed baseline).

`RefRasterize` mirrors `Rasterize.forward_gpu` / `backward_gpu` (reference rasterize.py:467-513, :849-889) call for
call: the same buffer allocation and initialisation, the same kernel sequence, the same compositing expressions --
with torch tensors for device memory and the reference's UNMODIFIED kernel strings, compiled ahead of time by
oracle/build_ref.py into oracle/_ref/nrref_<config>.so (launched with CuPy's geometry: 128-thread blocks,
ceil(n/128) blocks, grid-stride loop).  It is the exact parity oracle for face_index_map and the "reference CuPy
path timed on the same B200" baseline.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
from build_ref import OUT_DIR, config_key, normalize_config  # noqa: E402

_LIBS = {}


def lib_path(cfg):
    return os.path.join(OUT_DIR, "nrref_%s.so" % config_key(cfg))


def available(image_size, num_faces, texture_size=0, near=0.1, far=100, eps=1e-4, return_rgb=0, return_alpha=0,
              return_depth=0):
    return os.path.exists(lib_path(normalize_config(image_size, num_faces, texture_size, near, far, eps, return_rgb,
                                                    return_alpha, return_depth)))


def _load(cfg):
    path = lib_path(cfg)
    if path not in _LIBS:
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s missing: add the configuration to oracle/ref_configs.py and run oracle/build_ref.py where "
                "/root/reference exists (config %r)" % (path, cfg))
        _LIBS[path] = ctypes.CDLL(path)
    return _LIBS[path]


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class RefRasterize:
    def __init__(self, image_size, near, far, eps, background_color, return_rgb=False, return_alpha=False,
                 return_depth=False):
        if not any((return_rgb, return_alpha, return_depth)):
            raise Exception
        self.image_size, self.near, self.far, self.eps = image_size, near, far, eps
        self.background_color = background_color
        self.return_rgb, self.return_alpha, self.return_depth = bool(return_rgb), bool(return_alpha), bool(return_depth)

    def _call(self, name, *tensors_and_n):
        *tensors, n = tensors_and_n
        fn = getattr(self.lib, "launch_" + name)
        rc = fn(*[_p(t) for t in tensors], ctypes.c_longlong(n), _stream())
        if rc != 0:
            raise RuntimeError("reference kernel %s: cuda error %d" % (name, rc))

    # rasterize.py:467-513
    def forward(self, faces, textures=None):
        dev = faces.device
        self.faces = faces.detach().clone().contiguous()  # `inputs[0].copy()`
        bs, nf = self.faces.shape[:2]
        S = self.image_size
        self.batch_size, self.num_faces = bs, nf
        ts = 0
        if self.return_rgb:
            self.textures = textures.detach().contiguous()
            ts = self.texture_size = self.textures.shape[2]
        cfg = normalize_config(S, nf, ts, self.near, self.far, self.eps, self.return_rgb, self.return_alpha,
                               self.return_depth)
        self.lib = _load(cfg)
        f32, i32 = torch.float32, torch.int32
        self.face_index_map = -1 * torch.ones((bs, S, S), dtype=i32, device=dev)
        self.weight_map = torch.zeros((bs, S, S, 3), dtype=f32, device=dev)
        self.depth_map = torch.zeros((bs, S, S), dtype=f32, device=dev) + self.far
        one = lambda dt: torch.zeros(1, dtype=dt, device=dev)  # noqa: E731
        if self.return_rgb:
            self.rgb_map = torch.zeros((bs, S, S, 3), dtype=f32, device=dev)
            self.sampling_index_map = torch.zeros((bs, S, S, 8), dtype=i32, device=dev)
            self.sampling_weight_map = torch.zeros((bs, S, S, 8), dtype=f32, device=dev)
        else:
            self.rgb_map, self.sampling_index_map, self.sampling_weight_map = one(f32), one(i32), one(f32)
        self.alpha_map = torch.zeros((bs, S, S), dtype=f32, device=dev) if self.return_alpha else one(f32)
        self.face_inv_map = torch.zeros((bs, S, S, 3, 3), dtype=f32, device=dev) if self.return_depth else one(f32)

        # forward_face_index_map_gpu, safe path (rasterize.py:238-359)
        faces_inv = torch.zeros_like(self.faces)
        self._call("k1_face_inv", self.faces, faces_inv, bs * nf)
        self._call("k2_zbuffer", self.faces, faces_inv, self.face_index_map, self.weight_map, self.depth_map,
                   self.face_inv_map, bs * S * S)
        # forward_texture_sampling (:361-438)
        if self.return_rgb:
            self._call("k4_texture", self.faces, self.textures, self.face_index_map, self.weight_map, self.depth_map,
                       self.rgb_map, self.sampling_index_map, self.sampling_weight_map, bs * S * S)
            # forward_background_gpu (:451-465)
            bg = torch.as_tensor(self.background_color, dtype=f32, device=dev)
            mask = (0 <= self.face_index_map).to(f32)[:, :, :, None]
            if bg.dim() == 1:
                self.rgb_map = self.rgb_map * mask + (1 - mask) * bg[None, None, None, :]
            elif bg.dim() == 2:
                self.rgb_map = self.rgb_map * mask + (1 - mask) * bg[:, None, None, :]
        # forward_alpha_map_gpu (:440-449)
        if self.return_alpha:
            self.alpha_map[0 <= self.face_index_map] = 1
        rgb_r = self.rgb_map if self.return_rgb else None
        alpha_r = self.alpha_map.clone() if self.return_alpha else None
        depth_r = self.depth_map.clone() if self.return_depth else None
        return rgb_r, alpha_r, depth_r

    # rasterize.py:849-889
    def backward(self, grad_rgb=None, grad_alpha=None, grad_depth=None):
        dev = self.faces.device
        bs, nf, S = self.batch_size, self.num_faces, self.image_size
        f32 = torch.float32
        one = lambda: torch.zeros(1, dtype=f32, device=dev)  # noqa: E731
        self.grad_faces = torch.zeros_like(self.faces)
        self.grad_textures = torch.zeros_like(self.textures) if self.return_rgb else one()
        if self.return_rgb:
            g_rgb = grad_rgb.contiguous() if grad_rgb is not None else torch.zeros_like(self.rgb_map)
        else:
            g_rgb = one()
        if self.return_alpha:
            g_alpha = grad_alpha.contiguous() if grad_alpha is not None else torch.zeros_like(self.alpha_map)
        else:
            g_alpha = one()
        if self.return_depth:
            g_depth = grad_depth.contiguous() if grad_depth is not None else torch.zeros_like(self.depth_map)
        else:
            g_depth = one()
        rgb_map = self.rgb_map.contiguous()
        if self.return_rgb or self.return_alpha:
            self._call("k5_pixel_bwd", self.faces, self.face_index_map, rgb_map, self.alpha_map, g_rgb, g_alpha,
                       self.grad_faces, bs * nf)
        if self.return_rgb:
            self._call("k6_texture_bwd", self.face_index_map, self.sampling_weight_map, self.sampling_index_map,
                       g_rgb, self.grad_textures, bs * S * S)
        if self.return_depth:
            self._call("k7_depth_bwd", self.faces, self.depth_map, self.face_index_map, self.face_inv_map,
                       self.weight_map, g_depth, self.grad_faces, bs * S * S)
        return self.grad_faces, (self.grad_textures if self.return_rgb else None)


class RefResult(dict):
    pass


def rasterize_rgbad(faces, textures=None, image_size=256, anti_aliasing=True, near=0.1, far=100, eps=1e-4,
                    background_color=(0, 0, 0), return_rgb=True, return_alpha=True, return_depth=True):
    """rasterize.py:900-977 around the re-hosted kernels; `.backward(grad_rgb, grad_alpha, grad_depth)` takes
    gradients in API layout and returns (grad_faces, grad_textures)."""
    S = image_size * 2 if anti_aliasing else image_size
    fn = RefRasterize(S, near, far, eps, background_color, return_rgb, return_alpha, return_depth)
    rgb, alpha, depth = fn.forward(faces, textures)
    if return_rgb:
        rgb = rgb.permute(0, 3, 1, 2).flip(2)
    if return_alpha:
        alpha = alpha.flip(1)
    if return_depth:
        depth = depth.flip(1)
    if anti_aliasing:
        pool = torch.nn.functional.avg_pool2d
        if return_rgb:
            rgb = pool(rgb, 2, 2)
        if return_alpha:
            alpha = pool(alpha[:, None], 2, 2)[:, 0]
        if return_depth:
            depth = pool(depth[:, None], 2, 2)[:, 0]
    res = RefResult(rgb=rgb if return_rgb else None, alpha=alpha if return_alpha else None,
                    depth=depth if return_depth else None)
    res.fn = fn

    def backward(grad_rgb=None, grad_alpha=None, grad_depth=None):
        def back(g, is_rgb):
            if g is None:
                return None
            g = g.to(torch.float32)
            if anti_aliasing:
                g = g.repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1) * 0.25
            g = g.flip(-2)
            if is_rgb:
                g = g.permute(0, 2, 3, 1)
            return g.contiguous()
        return fn.backward(back(grad_rgb, True), back(grad_alpha, False), back(grad_depth, False))

    res.backward = backward
    return res
