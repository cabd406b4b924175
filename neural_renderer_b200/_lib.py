"""ctypes binding of libnr_b200.so -- the C ABI declared in include/nr_b200.h.

There is no CPU fallback and no other backend: if the CUDA library is missing the import of any hot-path entry
point fails loudly (the reference likewise raises NotImplementedError for CPU arrays, rasterize.py:893-897).
"""
from __future__ import annotations

import ctypes
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NR_B200_LIB", os.path.join(PKG_DIR, "libnr_b200.so"))  # override: kernel A/B experiments

NR_OK = 0
NR_RETURN_RGB = 1
NR_RETURN_ALPHA = 2
NR_RETURN_DEPTH = 4
NR_ANTI_ALIASING = 8
NR_BG_PER_BATCH = 16
NR_TEX_Z_BATCH0 = 32
NR_GRAD_ACCUMULATE = 64
NR_CAM_PERSPECTIVE = 0x100
NR_TEX_FILL_BACK = 0x400
NR_CAM_SHARED = 0x200
NR_FACES_INDEXED = 0x800
NR_INDICES_SHARED = 0x1000
NR_TEX_SHARED = 0x2000
NR_BWD_PART_TEXTURES = 0x4000
NR_BWD_PART_FACES = 0x8000
NR_FWD_STAGE_TEXTURES = 0x10000

ABI_VERSION = 3

# every symbol include/nr_b200.h declares
EXPORTED_SYMBOLS = (
    "nr_b200_abi_version",
    "nr_b200_error_string",
    "nr_b200_forward_workspace_bytes",
    "nr_b200_backward_workspace_bytes",
    "nr_b200_forward",
    "nr_b200_backward",
    "nr_b200_vertices_to_faces",
    "nr_b200_vertices_to_faces_backward",
    "nr_b200_camera_transform",
    "nr_b200_camera_transform_backward",
    "nr_b200_face_lighting",
    "nr_b200_face_lighting_backward",
    "nr_b200_bake_textures",
    "nr_b200_last_launch_count",
    "nr_b200_set_profiling",
    "nr_b200_read_profile",
)


class ForwardArgs(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("flags", ctypes.c_uint32),
        ("batch_size", ctypes.c_int32), ("num_faces", ctypes.c_int32),
        ("raster_size", ctypes.c_int32), ("texture_size", ctypes.c_int32),
        ("near_", ctypes.c_double), ("far_", ctypes.c_double), ("eps", ctypes.c_double),
        ("background", ctypes.c_float * 3), ("_pad0", ctypes.c_float),
        ("faces", ctypes.c_void_p), ("textures", ctypes.c_void_p), ("background_batch", ctypes.c_void_p),
        ("face_index_map", ctypes.c_void_p), ("weight_map", ctypes.c_void_p), ("depth_map", ctypes.c_void_p),
        ("rgb_map", ctypes.c_void_p), ("alpha_map", ctypes.c_void_p),
        ("out_rgb", ctypes.c_void_p), ("out_alpha", ctypes.c_void_p), ("out_depth", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("face_light", ctypes.c_void_p),
        ("vertices", ctypes.c_void_p), ("face_indices", ctypes.c_void_p),
        ("num_vertices", ctypes.c_int32), ("_pad1", ctypes.c_int32),
    ]


class BackwardArgs(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("flags", ctypes.c_uint32),
        ("batch_size", ctypes.c_int32), ("num_faces", ctypes.c_int32),
        ("raster_size", ctypes.c_int32), ("texture_size", ctypes.c_int32),
        ("eps", ctypes.c_double),
        ("faces", ctypes.c_void_p), ("textures", ctypes.c_void_p),
        ("face_index_map", ctypes.c_void_p), ("weight_map", ctypes.c_void_p), ("depth_map", ctypes.c_void_p),
        ("rgb_map", ctypes.c_void_p),
        ("grad_rgb", ctypes.c_void_p), ("grad_alpha", ctypes.c_void_p), ("grad_depth", ctypes.c_void_p),
        ("grad_faces", ctypes.c_void_p), ("grad_textures", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("face_light", ctypes.c_void_p), ("grad_face_light", ctypes.c_void_p),
        ("vertices", ctypes.c_void_p), ("face_indices", ctypes.c_void_p), ("grad_vertices", ctypes.c_void_p),
        ("num_vertices", ctypes.c_int32), ("_pad1", ctypes.c_int32),
    ]


_LIB = None


class LibraryMissing(ImportError):
    pass


def load():
    """dlopen libnr_b200.so; raises LibraryMissing (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            "%s not found: build it with `python -m neural_renderer_b200.build` (or __graft_entry__.build()); "
            "this package has no CPU or pure-PyTorch fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.nr_b200_abi_version.restype = ctypes.c_int
    lib.nr_b200_error_string.restype = ctypes.c_char_p
    lib.nr_b200_error_string.argtypes = [ctypes.c_int]
    for name in ("nr_b200_forward_workspace_bytes", "nr_b200_backward_workspace_bytes"):
        fn = getattr(lib, name)
        fn.restype = ctypes.c_size_t
        fn.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32]
    lib.nr_b200_forward.restype = ctypes.c_int
    lib.nr_b200_forward.argtypes = [ctypes.POINTER(ForwardArgs), ctypes.c_void_p]
    lib.nr_b200_backward.restype = ctypes.c_int
    lib.nr_b200_backward.argtypes = [ctypes.POINTER(BackwardArgs), ctypes.c_void_p]
    lib.nr_b200_vertices_to_faces.restype = ctypes.c_int
    lib.nr_b200_vertices_to_faces.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    lib.nr_b200_vertices_to_faces_backward.restype = ctypes.c_int
    lib.nr_b200_vertices_to_faces_backward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                                       ctypes.c_int32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
    lib.nr_b200_camera_transform.restype = ctypes.c_int
    lib.nr_b200_camera_transform.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32,
                                                                    ctypes.c_void_p, ctypes.c_void_p]
    lib.nr_b200_camera_transform_backward.restype = ctypes.c_int
    lib.nr_b200_camera_transform_backward.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32] + \
        [ctypes.c_void_p] * 5
    lib.nr_b200_face_lighting.restype = ctypes.c_int
    lib.nr_b200_face_lighting.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int32] * 3 + [ctypes.c_uint32, ctypes.c_void_p,
                                                                                       ctypes.c_void_p]
    lib.nr_b200_face_lighting_backward.restype = ctypes.c_int
    lib.nr_b200_face_lighting_backward.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int32] * 3 + [ctypes.c_uint32, ctypes.c_void_p,
                                                                                                ctypes.c_void_p]
    lib.nr_b200_bake_textures.restype = ctypes.c_int
    lib.nr_b200_bake_textures.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int32] * 4 + [ctypes.c_void_p, ctypes.c_void_p]
    lib.nr_b200_last_launch_count.restype = ctypes.c_int
    lib.nr_b200_set_profiling.restype = None
    lib.nr_b200_set_profiling.argtypes = [ctypes.c_int]
    lib.nr_b200_read_profile.restype = ctypes.c_int
    lib.nr_b200_read_profile.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
    if lib.nr_b200_abi_version() != ABI_VERSION:
        raise ImportError("libnr_b200.so ABI %d != binding ABI %d: rebuild" % (lib.nr_b200_abi_version(), ABI_VERSION))
    _LIB = lib
    return lib


def check(code):
    if code != NR_OK:
        raise RuntimeError("nr_b200: %s (code %d)" % (load().nr_b200_error_string(code).decode(), code))


def read_profile(max_entries=4096):
    """[(kernel name, milliseconds)] recorded since profiling was enabled / last read (synchronises)."""
    lib = load()
    names = ctypes.create_string_buffer(64 * max_entries)
    ms = (ctypes.c_float * max_entries)()
    n = lib.nr_b200_read_profile(names, len(names), ms, max_entries)
    parts = names.raw.split(b"\0")
    return [(parts[i].decode(), float(ms[i])) for i in range(n)]
