"""`Renderer` facade with the reference's attribute bag and three render methods (renderer.py:8-107)."""
from __future__ import annotations

import math

import torch

from . import functional as F
from .rasterize import rasterize, rasterize_depth, rasterize_silhouettes


class Renderer(object):
    def __init__(self):
        # rendering
        self.image_size = 256
        self.anti_aliasing = True
        self.background_color = [0, 0, 0]
        self.fill_back = True

        # camera
        self.perspective = True
        self.viewing_angle = 30
        self.eye = [0, 0, -(1. / math.tan(math.radians(self.viewing_angle)) + 1)]
        self.camera_mode = 'look_at'
        self.camera_direction = [0, 0, 1]
        self.near = 0.1
        self.far = 100

        # light
        self.light_intensity_ambient = 0.5
        self.light_intensity_directional = 0.5
        self.light_color_ambient = [1, 1, 1]  # white
        self.light_color_directional = [1, 1, 1]  # white
        self.light_direction = [0, 1, 0]  # up-to-down

        # rasterization
        self.rasterizer_eps = 1e-3

        # not in the reference: fold lighting / fill_back texture handling / vertices_to_faces into the rasterizer (same pixels)
        self.fused = True
        # rasterize.py:389: the reference samples the textures of EVERY batch item with the vertex depths of item 0.
        # None = module default (reference-exact, see neural_renderer_b200.set_reference_exact); False = every item with
        # its own depths (batches of different meshes / cameras, viewpoint shards of a multi-GPU run)
        self.reference_exact = None

    def _transform(self, vertices):
        # renderer.py:41-50 (look_at / look, then perspective), fused into one kernel on CUDA
        return F.camera_transform(vertices, self.eye, self.camera_mode, self.camera_direction, self.perspective,
                                  self.viewing_angle)

    @staticmethod
    def _fusable(vertices, faces):
        return vertices.is_cuda and faces.is_cuda and vertices.dtype == torch.float32 and not faces.is_floating_point()

    def _indices(self, faces):
        """Face indices as the rasterizer consumes them: a shared (expanded, stride-0) index set stays [1,F,3];
        fill_back appends the reversed copies (renderer.py:38-39) -- index data only, no vertex data is duplicated."""
        if faces.dim() == 3 and faces.shape[0] > 1 and faces.stride(0) == 0:
            faces = faces[:1]
        if self.fill_back:
            faces = torch.cat((faces, faces.flip(2)), dim=1)
        return faces

    def render_silhouettes(self, vertices, faces):
        if self.fused and self._fusable(vertices, faces):
            # vertices_to_faces (renderer.py:51) runs inside the rasterizer: no [B,F,3,3] tensor on either pass
            return rasterize_silhouettes(self._indices(faces), self.image_size, self.anti_aliasing,
                                         vertices=self._transform(vertices))
        if self.fill_back:
            faces = torch.cat((faces, faces.flip(2)), dim=1)
        vertices = self._transform(vertices)
        faces = F.vertices_to_faces(vertices, faces)
        # renderer.py:52 -- near / far / rasterizer_eps are NOT forwarded (module defaults apply)
        return rasterize_silhouettes(faces, self.image_size, self.anti_aliasing)

    def render_depth(self, vertices, faces):
        if self.fused and self._fusable(vertices, faces):
            return rasterize_depth(self._indices(faces), self.image_size, self.anti_aliasing,
                                   vertices=self._transform(vertices))
        if self.fill_back:
            faces = torch.cat((faces, faces.flip(2)), dim=1)
        vertices = self._transform(vertices)
        faces = F.vertices_to_faces(vertices, faces)
        return rasterize_depth(faces, self.image_size, self.anti_aliasing)  # renderer.py:72

    def render(self, vertices, faces, textures):
        fused = (self.fused and self._fusable(vertices, faces) and textures.is_cuda and textures.dtype == torch.float32)
        light_args = (self.light_intensity_ambient, self.light_intensity_directional, self.light_color_ambient,
                      self.light_color_directional, self.light_direction)
        if fused:
            # lighting.py:29-52, renderer.py:78-80 and vertices_to_faces (renderer.py:103) folded into the rasterizer:
            # neither `textures * light`, nor the doubled texture tensor, nor faces [B,F,3,3] exist; pixel values are
            # bit-identical to the op-by-op formulation
            indices = self._indices(faces)
            light = F.face_light_from_vertices(vertices, indices, *light_args)
            return rasterize(
                indices, textures, self.image_size, self.anti_aliasing, self.near, self.far, self.rasterizer_eps,
                self.background_color, face_light=light, textures_fill_back=self.fill_back,
                vertices=self._transform(vertices), reference_exact=self.reference_exact)
        if self.fill_back:
            faces = torch.cat((faces, faces.flip(2)), dim=1)
            textures = torch.cat((textures, textures.permute(0, 1, 4, 3, 2, 5)), dim=1)
        textures = F.lighting(F.vertices_to_faces(vertices, faces), textures, *light_args)
        vertices = self._transform(vertices)
        faces = F.vertices_to_faces(vertices, faces)
        return rasterize(
            faces, textures, self.image_size, self.anti_aliasing, self.near, self.far, self.rasterizer_eps,
            self.background_color, reference_exact=self.reference_exact)
