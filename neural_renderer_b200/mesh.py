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
        # mesh.py:29-34: broadcast for the minibatch.  The broadcasts are stride-0 views: `Renderer` keeps the index
        # and texture sets shared (NR_INDICES_SHARED / NR_TEX_SHARED) instead of materialising batch_size copies, and
        # the sigmoid runs once on the shared set (sigmoid(broadcast(x)) == broadcast(sigmoid(x))).
        vertices = self.vertices[None].expand(batch_size, *self.vertices.shape)
        faces = self.faces[None].expand(batch_size, *self.faces.shape)
        textures = torch.sigmoid(self.textures)[None].expand(batch_size, *self.textures.shape)
        return vertices, faces, textures

    def set_lr(self, lr_vertices, lr_textures):
        """mesh.py:36-38: per-parameter learning-rate multipliers honoured by `neural_renderer.Adam`."""
        self.vertices.lr = lr_vertices
        self.textures.lr = lr_textures
