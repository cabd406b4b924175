"""Thin array glue around the hot path, re-expressed with plain torch ops (SURVEY.md section 8(f) "next" rows; not
CUDA of their own yet).  Each function mirrors the reference function of the same name:

  cross                   cross.py:6-59
  get_points_from_angles  get_points_from_angles.py:6-24
  look_at                 look_at.py:7-46
  look                    look.py:7-45
  perspective             perspective.py:5-19   (pi is 3.1416 there, kept)
  lighting                lighting.py:8-52
  vertices_to_faces       vertices_to_faces.py:4-21
"""
from __future__ import annotations

import math

import torch


def _normalize(x, eps=1e-5):
    # chainer.functions.normalize (third-party, unpinned): x / (||x||_2 + eps) along axis 1
    return x / (x.norm(dim=1, keepdim=True) + eps)


def cross(a, b):
    """Row-wise 3-vector cross product of [N,3] arrays."""
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != 3 or b.shape[1] != 3 or a.shape[0] != b.shape[0]:
        raise ValueError("cross expects two [N,3] tensors")
    return torch.linalg.cross(a, b, dim=1)


def get_points_from_angles(distance, elevation, azimuth, degrees=True):
    if isinstance(distance, (float, int)):
        if degrees:
            elevation = math.radians(elevation)
            azimuth = math.radians(azimuth)
        return (
            distance * math.cos(elevation) * math.sin(azimuth),
            distance * math.sin(elevation),
            -distance * math.cos(elevation) * math.cos(azimuth))
    distance = torch.as_tensor(distance)
    elevation = torch.as_tensor(elevation, dtype=distance.dtype, device=distance.device)
    azimuth = torch.as_tensor(azimuth, dtype=distance.dtype, device=distance.device)
    if degrees:
        elevation = torch.deg2rad(elevation)
        azimuth = torch.deg2rad(azimuth)
    return torch.stack([
        distance * torch.cos(elevation) * torch.sin(azimuth),
        distance * torch.sin(elevation),
        -distance * torch.cos(elevation) * torch.cos(azimuth),
    ]).t()


def _as_vec(v, like):
    if isinstance(v, torch.Tensor):
        return v.to(device=like.device, dtype=like.dtype)
    return torch.tensor(v, dtype=like.dtype, device=like.device)


def look_at(vertices, eye, at=None, up=None):
    """"Look at" transformation of vertices [B,Nv,3]."""
    assert vertices.dim() == 3
    batch_size = vertices.shape[0]
    at = _as_vec([0, 0, 0] if at is None else at, vertices)
    up = _as_vec([0, 1, 0] if up is None else up, vertices)
    eye = _as_vec(eye, vertices)
    if eye.dim() == 1:
        eye = eye[None, :].expand(batch_size, 3)
    if at.dim() == 1:
        at = at[None, :].expand(batch_size, 3)
    if up.dim() == 1:
        up = up[None, :].expand(batch_size, 3)
    z_axis = _normalize(at - eye)
    x_axis = _normalize(cross(up, z_axis))
    y_axis = _normalize(cross(z_axis, x_axis))
    r = torch.stack((x_axis, y_axis, z_axis), dim=1)  # [bs,3,3]
    if r.shape[0] != vertices.shape[0]:
        r = r.expand(vertices.shape[0], 3, 3)
    vertices = vertices - eye[:, None, :]
    return torch.matmul(vertices, r.transpose(1, 2))


def look(vertices, eye, direction=None, up=None):
    """"Look" transformation of vertices [B,Nv,3] (camera at `eye` looking along `direction`)."""
    assert vertices.dim() == 3
    direction = _as_vec([0, 0, 1] if direction is None else direction, vertices)
    up = _as_vec([0, 1, 0] if up is None else up, vertices)
    eye = _as_vec(eye, vertices)
    if eye.dim() == 1:
        eye = eye[None, :]
    if direction.dim() == 1:
        direction = direction[None, :]
    if up.dim() == 1:
        up = up[None, :]
    z_axis = _normalize(direction)
    x_axis = _normalize(cross(up, z_axis))
    y_axis = _normalize(cross(z_axis, x_axis))
    r = torch.stack((x_axis, y_axis, z_axis), dim=1)
    if r.shape[0] != vertices.shape[0]:
        r = r.expand(vertices.shape[0], 3, 3)
    vertices = vertices - eye[:, None, :]
    return torch.matmul(vertices, r.transpose(1, 2))


def perspective(vertices, angle=30.):
    assert vertices.dim() == 3
    if isinstance(angle, (float, int)):
        angle = torch.tensor(float(angle), dtype=vertices.dtype, device=vertices.device)
    angle = angle / 180. * 3.1416
    angle = angle[None].expand(vertices.shape[0]) if angle.dim() == 0 else angle
    width = torch.tan(angle)[:, None]
    z = vertices[:, :, 2]
    x = vertices[:, :, 0] / z / width
    y = vertices[:, :, 1] / z / width
    return torch.stack((x, y, z), dim=2)


def lighting(faces, textures, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
             color_directional=(1, 1, 1), direction=(0, 1, 0)):
    bs, nf = faces.shape[:2]
    color_ambient = _as_vec(color_ambient, faces)
    color_directional = _as_vec(color_directional, faces)
    direction = _as_vec(direction, faces)
    if color_ambient.dim() == 1:
        color_ambient = color_ambient[None, :].expand(bs, 3)
    if color_directional.dim() == 1:
        color_directional = color_directional[None, :].expand(bs, 3)
    if direction.dim() == 1:
        direction = direction[None, :].expand(bs, 3)
    light = torch.zeros((bs, nf, 3), dtype=faces.dtype, device=faces.device)
    if intensity_ambient != 0:
        light = light + intensity_ambient * color_ambient[:, None, :]
    if intensity_directional != 0:
        f = faces.reshape(bs * nf, 3, 3)
        v10 = f[:, 0] - f[:, 1]
        v12 = f[:, 2] - f[:, 1]
        normals = _normalize(cross(v10, v12)).reshape(bs, nf, 3)
        cos = torch.relu((normals * direction[:, None, :]).sum(dim=2))
        light = light + intensity_directional * color_directional[:, None, :] * cos[:, :, None]
    return textures * light[:, :, None, None, None, :]


class _VerticesToFaces(torch.autograd.Function):
    """CUDA gather (forward) / scatter-add (backward) behind the C ABI (nr_b200_vertices_to_faces*)."""

    @staticmethod
    def forward(ctx, vertices, faces_i32):
        import ctypes
        from . import _lib
        lib = _lib.load()
        v = vertices.detach().contiguous()
        bs, nv = v.shape[:2]
        nf = faces_i32.shape[1]
        out = torch.empty((bs, nf, 3, 3), dtype=torch.float32, device=v.device)
        with torch.cuda.device(v.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(v.device).cuda_stream)
            _lib.check(lib.nr_b200_vertices_to_faces(v.data_ptr(), faces_i32.data_ptr(), bs, nv, nf, out.data_ptr(), stream))
        ctx.save_for_backward(faces_i32)
        ctx.nv = nv
        return out

    @staticmethod
    def backward(ctx, grad_out):
        import ctypes
        from . import _lib
        lib = _lib.load()
        faces_i32, = ctx.saved_tensors
        g = grad_out.detach().to(torch.float32).contiguous()
        bs, nf = g.shape[:2]
        grad_v = torch.empty((bs, ctx.nv, 3), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
            _lib.check(lib.nr_b200_vertices_to_faces_backward(g.data_ptr(), faces_i32.data_ptr(), bs, ctx.nv, nf,
                                                              grad_v.data_ptr(), 0, stream))
        return grad_v, None


def vertices_to_faces(vertices, faces):
    """[B,Nv,3] x [B,Nf,3] int -> [B,Nf,3,3]"""
    assert vertices.dim() == 3
    assert faces.dim() == 3
    assert vertices.shape[0] == faces.shape[0]
    assert vertices.shape[2] == 3
    assert faces.shape[2] == 3
    bs, nv = vertices.shape[:2]
    if vertices.is_cuda and faces.is_cuda and vertices.dtype == torch.float32 and bs <= 65535:
        return _VerticesToFaces.apply(vertices, faces.to(torch.int32).contiguous())
    idx = faces.long() + (torch.arange(bs, device=faces.device, dtype=torch.long) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[idx]
