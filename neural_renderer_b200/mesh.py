"""`Mesh` parameter holder (mesh.py:7-38) as a torch.nn.Module."""
from __future__ import annotations

import torch

from .io import load_obj


class Mesh(torch.nn.Module):
    def __init__(self, filename_obj, texture_size=4, normalization=True):
        super().__init__()
        vertices, faces = load_obj(filename_obj, normalization)
        self.vertices = torch.nn.Parameter(torch.from_numpy(vertices))
        self.register_buffer('faces', torch.from_numpy(faces))
        self.num_vertices = self.vertices.shape[0]
        self.num_faces = self.faces.shape[0]
        shape = (self.num_faces, texture_size, texture_size, texture_size, 3)
        self.textures = torch.nn.Parameter(torch.randn(shape) * 0.05)  # chainer.initializers.Normal() scale
        self.texture_size = texture_size

    def get_batch(self, batch_size):
        vertices = self.vertices[None].expand(batch_size, *self.vertices.shape)
        faces = self.faces[None].expand(batch_size, *self.faces.shape)
        textures = torch.sigmoid(self.textures[None].expand(batch_size, *self.textures.shape))
        return vertices, faces, textures
