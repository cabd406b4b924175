"""The steps either side of the rasterizer, each mirroring the reference function of the same
name.  On CUDA float32 tensors the camera pipeline, the per-face light factor and vertices_to_faces run as fused
kernels behind the C ABI (csrc/nr_glue.cu); the plain torch formulation below them serves CPU tensors and is the
test oracle of those kernels.  `Renderer` goes one step further and hands vertices + indices to the rasterizer itself
(NR_FACES_INDEXED), so vertices_to_faces does not run at all on its path.

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


class _CameraTransform(torch.autograd.Function):
    """out = perspective(rot @ (vertices - eye)) as one CUDA kernel each way (nr_b200_camera_transform*).

    rot [C,3,3] / eye [C,3] / width [C] with C == batch or C == 1 (one camera for every item); rot or eye may be
    None (identity / origin); width None = no perspective division."""

    @staticmethod
    def forward(ctx, vertices, rot, eye, width):
        import ctypes
        from . import _lib
        lib = _lib.load()
        v = vertices.detach().contiguous()
        bs, nv = v.shape[:2]
        cams = [t.shape[0] for t in (rot, eye, width) if t is not None]
        shared = all(c == 1 for c in cams) and bs != 1
        if any(c != (1 if shared else bs) for c in cams):
            raise ValueError("camera arrays must hold one item per batch entry or exactly one")
        flags = (_lib.NR_CAM_SHARED if shared else 0) | (_lib.NR_CAM_PERSPECTIVE if width is not None else 0)
        r = None if rot is None else rot.detach().to(torch.float32).contiguous()
        e = None if eye is None else eye.detach().to(torch.float32).contiguous()
        w = None if width is None else width.detach().to(torch.float32).contiguous()
        out = torch.empty_like(v)
        ptr = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(v.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(v.device).cuda_stream)
            _lib.check(lib.nr_b200_camera_transform(v.data_ptr(), ptr(r), ptr(e), ptr(w), bs, nv, flags, out.data_ptr(), stream))
        ctx.save_for_backward(v, r, e, w)
        ctx.flags = flags
        return out

    @staticmethod
    def backward(ctx, grad_out):
        import ctypes
        from . import _lib
        lib = _lib.load()
        v, r, e, w = ctx.saved_tensors
        g = grad_out.detach().to(torch.float32).contiguous()
        bs, nv = v.shape[:2]
        need_v, need_r, need_e, need_w = ctx.needs_input_grad
        gv = torch.empty_like(v) if need_v else None
        gr = torch.empty_like(r) if (need_r and r is not None) else None
        ge = torch.empty_like(e) if (need_e and e is not None) else None
        gw = torch.empty_like(w) if (need_w and w is not None) else None
        ptr = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(v.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(v.device).cuda_stream)
            _lib.check(lib.nr_b200_camera_transform_backward(v.data_ptr(), ptr(r), ptr(e), ptr(w), g.data_ptr(), bs, nv,
                                                             ctx.flags, ptr(gv), ptr(gr), ptr(ge), ptr(gw), stream))
        return gv, gr, ge, gw


def _fused_camera_ok(vertices):
    return vertices.is_cuda and vertices.dtype == torch.float32 and vertices.shape[0] <= 65535


_CAMERA_CACHE = {}  # small device constants (camera frames, tan(angle), light parameters) keyed by their host values


def _cache_put(key, value):
    if len(_CAMERA_CACHE) >= 256:  # bounded: a caller that animates the camera / light with Python floats must not leak
        _CAMERA_CACHE.clear()
    _CAMERA_CACHE[key] = value
    return value


def _host_camera(kind, eye, aux, up, device):
    """Rotation / eye of a camera given as plain Python numbers: evaluated once on the host in float32 with the same
    operation order as the tensor path, cached per (camera, device) -- zero device work per call."""
    import numpy as np
    key = (kind, tuple(float(x) for x in eye), tuple(float(x) for x in aux), tuple(float(x) for x in up), str(device))
    hit = _CAMERA_CACHE.get(key)
    if hit is not None:
        return hit
    f32 = np.float32

    def normalize(x):
        return x / (np.sqrt((x * x).sum(dtype=f32), dtype=f32) + f32(1e-5))
    e = np.asarray(key[1], dtype=f32)
    a = np.asarray(key[2], dtype=f32)
    u = np.asarray(key[3], dtype=f32)
    z_axis = normalize((a - e) if kind == "look_at" else a)
    x_axis = normalize(np.cross(u, z_axis).astype(f32))
    y_axis = normalize(np.cross(z_axis, x_axis).astype(f32))
    rot = torch.from_numpy(np.stack((x_axis, y_axis, z_axis))[None].astype(f32)).to(device)
    eye_t = torch.from_numpy(e[None]).to(device)
    return _cache_put(key, (rot, eye_t))


def _is_plain(*vals):
    return all(v is None or not isinstance(v, torch.Tensor) for v in vals)


def _camera_frame(kind, vertices, eye, aux, up):
    """(rot [C,3,3], eye [C,3]) of look_at (aux = at) / look (aux = direction); C = 1 or batch."""
    batch_size = vertices.shape[0]
    aux_default = [0, 0, 0] if kind == "look_at" else [0, 0, 1]
    if _is_plain(eye, aux, up) and _fused_camera_ok(vertices):
        import numpy as np
        eye_np = np.asarray(eye, dtype=np.float64)
        if eye_np.ndim == 1:
            return _host_camera(kind, eye_np, aux_default if aux is None else aux, [0, 1, 0] if up is None else up,
                                vertices.device)
    aux = _as_vec(aux_default if aux is None else aux, vertices)
    up = _as_vec([0, 1, 0] if up is None else up, vertices)
    eye = _as_vec(eye, vertices)
    if eye.dim() == 1:
        eye = eye[None, :]
    if aux.dim() == 1:
        aux = aux[None, :]
    if up.dim() == 1:
        up = up[None, :]
    n = max(eye.shape[0], aux.shape[0], up.shape[0])
    z_axis = _normalize((aux - eye) if kind == "look_at" else aux)
    if z_axis.shape[0] != n:
        z_axis = z_axis.expand(n, 3)
    x_axis = _normalize(cross(up.expand(n, 3), z_axis))
    y_axis = _normalize(cross(z_axis, x_axis))
    rot = torch.stack((x_axis, y_axis, z_axis), dim=1)  # [C,3,3]
    if n not in (1, batch_size):
        raise ValueError("camera batch does not match the vertices")
    return rot, (eye if eye.shape[0] == n else eye.expand(n, 3))


def _perspective_width(vertices, angle):
    """tan(angle / 180 * 3.1416) as a [C] tensor (perspective.py:12-14; pi is 3.1416 there, kept)."""
    if isinstance(angle, (float, int)):
        key = ("width", float(angle), str(vertices.device), str(vertices.dtype))
        hit = _CAMERA_CACHE.get(key)
        if hit is None:
            a = torch.tensor(float(angle), dtype=vertices.dtype, device=vertices.device) / 180. * 3.1416
            hit = _cache_put(key, torch.tan(a)[None])
        return hit
    angle = angle / 180. * 3.1416
    angle = angle[None] if angle.dim() == 0 else angle
    return torch.tan(angle)


def camera_transform(vertices, eye, camera_mode="look_at", camera_direction=None, perspective=True, viewing_angle=30.,
                     at=None, up=None):
    """Renderer's camera pipeline (renderer.py:41-50): look_at / look, then perspective -- one fused kernel each way on
    CUDA float32 vertices, the reference's op-by-op formulation otherwise."""
    assert vertices.dim() == 3
    if not _fused_camera_ok(vertices):
        if camera_mode == "look_at":
            vertices = look_at(vertices, eye, at, up)
        elif camera_mode == "look":
            vertices = look(vertices, eye, camera_direction, up)
        return perspective_(vertices, viewing_angle) if perspective else vertices
    rot = eye_t = None
    if camera_mode == "look_at":
        rot, eye_t = _camera_frame("look_at", vertices, eye, at, up)
    elif camera_mode == "look":
        rot, eye_t = _camera_frame("look", vertices, eye, camera_direction, up)
    width = _perspective_width(vertices, viewing_angle) if perspective else None
    if rot is None and width is None:
        return vertices
    if rot is not None and width is not None and rot.shape[0] != width.shape[0]:
        n = max(rot.shape[0], width.shape[0])
        rot, eye_t, width = rot.expand(n, 3, 3), eye_t.expand(n, 3), width.expand(n)
    return _CameraTransform.apply(vertices, rot, eye_t, width)


def look_at(vertices, eye, at=None, up=None):
    """"Look at" transformation of vertices [B,Nv,3]."""
    if vertices.dim() == 3 and _fused_camera_ok(vertices):
        rot, eye_t = _camera_frame("look_at", vertices, eye, at, up)
        return _CameraTransform.apply(vertices, rot, eye_t, None)
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
    if vertices.dim() == 3 and _fused_camera_ok(vertices):
        rot, eye_t = _camera_frame("look", vertices, eye, direction, up)
        return _CameraTransform.apply(vertices, rot, eye_t, None)
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
    if _fused_camera_ok(vertices):
        return _CameraTransform.apply(vertices, None, None, _perspective_width(vertices, angle))
    if isinstance(angle, (float, int)):
        angle = torch.tensor(float(angle), dtype=vertices.dtype, device=vertices.device)
    angle = angle / 180. * 3.1416
    angle = angle[None].expand(vertices.shape[0]) if angle.dim() == 0 else angle
    width = torch.tan(angle)[:, None]
    z = vertices[:, :, 2]
    x = vertices[:, :, 0] / z / width
    y = vertices[:, :, 1] / z / width
    return torch.stack((x, y, z), dim=2)


perspective_ = perspective


def face_light(faces, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
               color_directional=(1, 1, 1), direction=(0, 1, 0)):
    """Per-face RGB light factor [B,F,3] of lighting.py:29-51 (ambient + Lambertian directional)."""
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
    return light


class _FaceLighting(torch.autograd.Function):
    """face_light [B,F,3] straight from vertices and face indices (nr_b200_face_lighting*); params [C,9], C in {1, B}."""

    @staticmethod
    def forward(ctx, vertices, faces_i32, params):
        import ctypes
        from . import _lib
        lib = _lib.load()
        v = vertices.detach().contiguous()
        bs, nv = v.shape[:2]
        nf = faces_i32.shape[1]
        flags = _lib.NR_CAM_SHARED if (params.shape[0] == 1 and bs != 1) else 0
        if faces_i32.shape[0] == 1 and bs != 1:
            flags |= _lib.NR_INDICES_SHARED  # one index set for every batch item
        out = torch.empty((bs, nf, 3), dtype=torch.float32, device=v.device)
        with torch.cuda.device(v.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(v.device).cuda_stream)
            _lib.check(lib.nr_b200_face_lighting(v.data_ptr(), faces_i32.data_ptr(), params.data_ptr(), bs, nv, nf, flags,
                                                 out.data_ptr(), stream))
        ctx.save_for_backward(v, faces_i32, params)
        ctx.flags = flags
        return out

    @staticmethod
    def backward(ctx, grad_light):
        import ctypes
        from . import _lib
        lib = _lib.load()
        v, faces_i32, params = ctx.saved_tensors
        g = grad_light.detach().to(torch.float32).contiguous()
        bs, nv = v.shape[:2]
        grad_v = torch.empty_like(v)
        with torch.cuda.device(v.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(v.device).cuda_stream)
            _lib.check(lib.nr_b200_face_lighting_backward(v.data_ptr(), faces_i32.data_ptr(), params.data_ptr(), g.data_ptr(), bs,
                                                          nv, faces_i32.shape[1], ctx.flags, grad_v.data_ptr(), stream))
        return grad_v, None, None


def face_light_from_vertices(vertices, faces, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
                             color_directional=(1, 1, 1), direction=(0, 1, 0)):
    """`face_light(vertices_to_faces(vertices, faces), ...)` without the gathered tensor: one kernel each way on CUDA."""
    plain = _is_plain(intensity_ambient, intensity_directional, color_ambient, color_directional, direction)
    if not (plain and _fused_camera_ok(vertices) and faces.is_cuda):
        return face_light(vertices_to_faces(vertices, faces), intensity_ambient, intensity_directional, color_ambient,
                          color_directional, direction)
    import numpy as np
    ca, cd, d = (np.asarray(x, dtype=np.float32) for x in (color_ambient, color_directional, direction))
    if ca.ndim != 1 or cd.ndim != 1 or d.ndim != 1:
        return face_light(vertices_to_faces(vertices, faces), intensity_ambient, intensity_directional, color_ambient,
                          color_directional, direction)
    key = ("light", float(intensity_ambient), float(intensity_directional), tuple(ca.tolist()), tuple(cd.tolist()),
           tuple(d.tolist()), str(vertices.device))
    params = _CAMERA_CACHE.get(key)
    if params is None:
        row = np.concatenate([np.float32(intensity_ambient) * ca, np.float32(intensity_directional) * cd, d]).astype(np.float32)
        params = _cache_put(key, torch.from_numpy(row[None]).to(vertices.device))
    if faces.dim() == 3 and faces.shape[0] > 1 and faces.stride(0) == 0:
        faces = faces[:1]  # expanded shared index set: keep it shared
    return _FaceLighting.apply(vertices, faces.to(torch.int32).contiguous(), params)


def lighting(faces, textures, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
             color_directional=(1, 1, 1), direction=(0, 1, 0)):
    light = face_light(faces, intensity_ambient, intensity_directional, color_ambient, color_directional, direction)
    return textures * light[:, :, None, None, None, :]  # lighting.py:52


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
