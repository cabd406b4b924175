"""OBJ I/O with the reference's conventions (load_obj.py:8-197, save_obj.py:151-191).

File parsing is host-side numpy; the one compute step -- baking a texture image into the per-face ts^3 cubes
(load_obj.py:88-137, a CuPy kernel in the reference) -- runs as `nr_b200_bake_textures` on the GPU, like the
reference needs a GPU for `load_texture=True`.  Results are numpy arrays, as in the reference."""
from __future__ import annotations

import os

import numpy as np


def load_mtl(filename_mtl):
    """Kd colours and map_Kd texture file names per material (load_obj.py:9-22); file order is kept."""
    texture_filenames = {}
    colors = {}
    material_name = ''
    with open(filename_mtl) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == 'newmtl':
                material_name = tok[1]
            if tok[0] == 'map_Kd':
                texture_filenames[material_name] = tok[1]
            if tok[0] == 'Kd':
                colors[material_name] = np.array([float(v) for v in tok[1:4]], dtype=np.float32)
    return colors, texture_filenames


def _read_image(path):
    """RGB image as float32 in [0, 1] (skimage.io.imread(...) / 255 in the reference, load_obj.py:80)."""
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert('RGB'), dtype=np.float32) / np.float32(255.)


def parse_texture_faces(filename_obj):
    """UV triangle per face [F,3,2] and the material name of every face (load_obj.py:27-64)."""
    vt = []
    with open(filename_obj) as f:
        lines = f.readlines()
    for line in lines:
        tok = line.split()
        if tok and tok[0] == 'vt':
            vt.append([float(v) for v in tok[1:3]])
    vt = np.array(vt, dtype=np.float32).reshape(-1, 2)
    faces, material_names, material_name = [], [], ''
    for line in lines:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == 'f':
            vs = tok[1:]
            idx = [int(v.split('/')[1]) if ('/' in v and v.split('/')[1] != '') else 0 for v in vs]
            for i in range(len(vs) - 2):
                faces.append((idx[0], idx[i + 1], idx[i + 2]))
                material_names.append(material_name)
        if tok[0] == 'usemtl':
            material_name = tok[1]
    faces = np.array(faces, dtype=np.int32).reshape(-1, 3) - 1  # a missing index becomes -1 = the last vt, as in the reference
    if vt.shape[0] == 0:
        vt = np.zeros((1, 2), dtype=np.float32)
    uv = vt[faces]                                               # [F,3,2]
    wrap = uv > 1
    uv[wrap] = uv[wrap] % 1                                      # load_obj.py:66
    return np.ascontiguousarray(uv, dtype=np.float32), material_names


def bake_textures(image, uv_faces, is_update, texture_size, textures):
    """The bilinear bake of load_obj.py:88-137 on the GPU (`image` [H,W,3] float32, rows already flipped)."""
    import ctypes

    import torch
    from . import _lib
    if not torch.cuda.is_available():
        raise NotImplementedError("texture baking runs on the GPU (the reference's load_textures needs one as well)")
    lib = _lib.load()
    dev = torch.device("cuda")
    img = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(dev)
    uv = torch.from_numpy(np.ascontiguousarray(uv_faces, dtype=np.float32)).to(dev)
    upd = None if is_update is None else torch.from_numpy(np.ascontiguousarray(is_update, dtype=np.int32)).to(dev)
    tex = torch.from_numpy(np.ascontiguousarray(textures, dtype=np.float32)).to(dev)
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.nr_b200_bake_textures(img.data_ptr(), uv.data_ptr(), None if upd is None else upd.data_ptr(),
                                             uv.shape[0], int(texture_size), img.shape[0], img.shape[1], tex.data_ptr(),
                                             stream))
    return tex.cpu().numpy()


def load_textures(filename_obj, filename_mtl, texture_size):
    """Per-face texture cubes [F,ts,ts,ts,3] from the OBJ's UVs and its MTL (load_obj.py:25-144): 0.5 grey, then the
    material's Kd colour, then -- for materials with a map_Kd image -- the bilinear bake."""
    uv_faces, material_names = parse_texture_faces(filename_obj)
    colors, texture_filenames = load_mtl(filename_mtl)
    nf = uv_faces.shape[0]
    textures = np.zeros((nf, texture_size, texture_size, texture_size, 3), dtype=np.float32) + np.float32(0.5)
    names = np.array(material_names)
    for material_name, color in colors.items():
        textures[names == material_name] = color[None, None, None, None, :]
    for material_name, filename_texture in texture_filenames.items():
        path = os.path.join(os.path.dirname(filename_obj), filename_texture)
        image = _read_image(path)[::-1, ::1]                    # load_obj.py:82
        is_update = (names == material_name).astype(np.int32)
        textures = bake_textures(image, uv_faces, is_update, texture_size, textures)
    return textures


def load_obj(filename_obj, normalization=True, texture_size=4, load_texture=False):
    """Load vertices (v x y z), faces (f ...; n-gons are fan-triangulated) and optionally the baked per-face textures
    of a Wavefront .obj file (load_obj.py:147-197)."""
    vertices, faces = [], []
    with open(filename_obj) as f:
        lines = f.readlines()
    for line in lines:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == 'v':
            vertices.append([float(v) for v in tok[1:4]])
        elif tok[0] == 'f':
            vs = [int(t.split('/')[0]) for t in tok[1:]]
            for i in range(len(vs) - 2):
                faces.append((vs[0], vs[i + 1], vs[i + 2]))
    vertices = np.array(vertices, dtype=np.float32)
    faces = np.array(faces, dtype=np.int32) - 1
    textures = None
    if load_texture:
        for line in lines:
            if line.startswith('mtllib'):
                filename_mtl = os.path.join(os.path.dirname(filename_obj), line.split()[1])
                textures = load_textures(filename_obj, filename_mtl, texture_size)
        if textures is None:
            raise Exception('Failed to load textures.')  # load_obj.py:185
    if normalization:  # unit cube centred at zero, load_obj.py:188-192
        vertices -= vertices.min(0)[None, :]
        vertices /= np.abs(vertices).max()
        vertices *= 2
        vertices -= vertices.max(0)[None, :] / 2
    if load_texture:
        return vertices, faces, textures
    return vertices, faces


def save_obj(filename, vertices, faces, textures=None):
    """Write geometry (v / f lines).  Texture atlas export (save_obj.py:10-148) is outside the hot-path scope."""
    if textures is not None:
        raise NotImplementedError("texture atlas export is outside the B200 hot path")
    vertices = np.asarray(vertices)
    faces = np.asarray(faces)
    with open(filename, 'w') as f:
        for v in vertices:
            f.write('v %.8f %.8f %.8f\n' % (v[0], v[1], v[2]))
        f.write('\n')
        for face in faces:
            f.write('f %d %d %d\n' % (face[0] + 1, face[1] + 1, face[2] + 1))
