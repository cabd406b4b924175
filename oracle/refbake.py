"""Host for the reference's texture-bake kernel string (load_obj.py:88-137), compiled by build_ref.build_bake into
oracle/_ref/ (TEST INFRASTRUCTURE -- only tests may import this).  Mirrors the call at load_obj.py:138-144."""
import ctypes
import os

import torch

import build_ref

_LIBS = {}


def available(texture_size, image_height, image_width):
    return os.path.exists(build_ref.bake_lib_path(texture_size, image_height, image_width))


def bake(image, uv_faces, is_update, texture_size, textures):
    """image [H,W,3] (already flipped), uv_faces [F,3,2], is_update [F] int32, textures [F,ts,ts,ts,3]: CUDA tensors.
    The kernel reads one row / column past the image when a coordinate is exactly 1 (with weight 0): the image is
    staged with a zero row behind it so that those reads are defined."""
    H, W = int(image.shape[0]), int(image.shape[1])
    path = build_ref.bake_lib_path(texture_size, H, W)
    lib = _LIBS.get(path)
    if lib is None:
        lib = _LIBS[path] = ctypes.CDLL(path)
    padded = torch.zeros(((H + 2) * W, 3), dtype=torch.float32, device=image.device)
    padded[:H * W] = image.reshape(H * W, 3)
    out = textures.clone().contiguous()
    uv = uv_faces.contiguous().float()
    upd = is_update.contiguous().to(torch.int32)
    n = out.numel() // 3
    stream = ctypes.c_void_p(torch.cuda.current_stream(image.device).cuda_stream)
    rc = lib.launch_k_bake(ctypes.c_void_p(padded.data_ptr()), ctypes.c_void_p(uv.data_ptr()),
                           ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(upd.data_ptr()), ctypes.c_longlong(n), stream)
    if rc:
        raise RuntimeError("reference bake kernel launch failed: %d" % rc)
    return out
