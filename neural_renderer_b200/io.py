"""Minimal OBJ I/O so the reference's examples run (load_obj.py:147-197, save_obj.py:151-191).
File parsing is host-side, runs once and is out of the hot-path scope: geometry only (no MTL/texture baking)."""
from __future__ import annotations

import numpy as np


def load_obj(filename_obj, normalization=True, texture_size=4, load_texture=False):
    """Load vertices (v x y z) and faces (f ...; n-gons are fan-triangulated) of a Wavefront .obj file."""
    if load_texture:
        raise NotImplementedError("texture baking from MTL/images (load_obj.py:25-144) is outside the B200 hot path")
    vertices, faces = [], []
    with open(filename_obj) as f:
        for line in f:
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
    if normalization:  # unit cube centred at zero, load_obj.py:188-192
        vertices -= vertices.min(0)[None, :]
        vertices /= np.abs(vertices).max()
        vertices *= 2
        vertices -= vertices.max(0)[None, :] / 2
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
