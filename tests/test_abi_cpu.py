"""CPU: the C-ABI library loads and exports every symbol include/nr_b200.h declares (no compute without a GPU);
host-side argument checking; the reference's export list is importable."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from neural_renderer_b200 import build, _lib
    build.build_library()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from neural_renderer_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "nr_b200.h")).read()
    declared = set(re.findall(r"\b(nr_b200_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_struct_sizes(lib):
    from neural_renderer_b200 import _lib
    assert lib.nr_b200_abi_version() == _lib.ABI_VERSION
    # a wrong struct_size must be rejected before anything touches the device
    a = _lib.ForwardArgs()
    a.struct_size = 4
    assert lib.nr_b200_forward(ctypes.byref(a), None) == -1
    b = _lib.BackwardArgs()
    b.struct_size = 4
    assert lib.nr_b200_backward(ctypes.byref(b), None) == -1
    assert b"workspace" in lib.nr_b200_error_string(-2)


def test_workspace_queries_are_pure_host(lib):
    n = lib.nr_b200_forward_workspace_bytes(64, 5000, 256, 4, 1)
    assert n >= 64 * 5000 * 8
    assert lib.nr_b200_backward_workspace_bytes(64, 5000, 256, 4, 1) >= 64 * 5000 * 8
    assert lib.nr_b200_forward_workspace_bytes(0, 0, 0, 0, 0) >= 16


def test_invalid_arguments_rejected_on_host(lib):
    from neural_renderer_b200 import _lib
    a = _lib.ForwardArgs()
    a.struct_size = ctypes.sizeof(_lib.ForwardArgs)
    a.batch_size, a.num_faces, a.raster_size = 1, 1, 8
    a.flags = 0  # nothing to draw (rasterize.py:25-27)
    assert lib.nr_b200_forward(ctypes.byref(a), None) == -1


def test_export_list_matches_reference():
    import neural_renderer
    for name in ("cross get_points_from_angles lighting load_obj look look_at Mesh Adam perspective rasterize_rgbad "
                 "rasterize rasterize_silhouettes rasterize_depth use_unsafe_rasterizer Rasterize Renderer save_obj "
                 "vertices_to_faces __version__").split():
        assert hasattr(neural_renderer, name), name


def test_host_side_errors():
    import neural_renderer as nr
    faces = torch.zeros(1, 2, 3, 3)
    with pytest.raises(NotImplementedError):  # no CPU path (reference: forward_cpu raises, rasterize.py:893-897)
        nr.rasterize_silhouettes(faces, 32)
    with pytest.raises(ValueError):
        nr.rasterize_silhouettes(torch.zeros(1, 2, 3, 2), 32)
    with pytest.raises(TypeError):
        nr.rasterize_silhouettes(torch.zeros(1, 2, 3, 3, dtype=torch.int32), 32)
    with pytest.raises(Exception):
        nr.Rasterize(32, 0.1, 100, 1e-4, (0, 0, 0))  # nothing to draw
    with pytest.raises((TypeError, ValueError)):
        nr.rasterize(faces, None, 32)


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "neural_renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "nr_oracle" not in src and "refhost" not in src and "oracle/" not in src, f


def test_glue_entry_points_reject_null_arguments(lib):
    # every glue entry point validates its arguments on the host before any launch
    assert lib.nr_b200_vertices_to_faces(None, None, 1, 1, 1, None, None) == -1
    assert lib.nr_b200_camera_transform(None, None, None, None, 1, 1, 0, None, None) == -1
    assert lib.nr_b200_camera_transform_backward(None, None, None, None, None, 1, 1, 0, None, None, None, None, None) == -1
    assert lib.nr_b200_face_lighting(None, None, None, 1, 1, 1, 0, None, None) == -1
    assert lib.nr_b200_face_lighting_backward(None, None, None, None, 1, 1, 1, 0, None, None) == -1
    assert lib.nr_b200_bake_textures(None, None, None, 1, 4, 8, 8, None, None) == -1


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of the two argument structs as a C compiler sees include/nr_b200.h == the ctypes mirror."""
    import subprocess
    from neural_renderer_b200 import _lib
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nr_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(nr_b200_forward_args), offsetof(nr_b200_forward_args, faces), offsetof(nr_b200_forward_args, face_light),'
                   'sizeof(nr_b200_backward_args), offsetof(nr_b200_backward_args, faces), offsetof(nr_b200_backward_args, grad_face_light));'
                   'return 0;}\n')
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.ForwardArgs), _lib.ForwardArgs.faces.offset, _lib.ForwardArgs.face_light.offset,
            ctypes.sizeof(_lib.BackwardArgs), _lib.BackwardArgs.faces.offset, _lib.BackwardArgs.grad_face_light.offset]
    assert got == want


def test_no_environment_switches_in_the_product_build():
    """Ablation / tuning knobs exist only behind NR_B200_DEBUG_KNOBS / NR_B200_TUNING (off in the product build): a
    stray environment variable must not be able to change what the library computes."""
    import re
    csrc = os.path.join(ROOT, "neural_renderer_b200", "csrc")
    for name in os.listdir(csrc):
        depth, guarded = [], 0
        for line in open(os.path.join(csrc, name)):
            s = line.strip()
            if re.match(r"#\s*if", s):
                depth.append(bool(re.search(r"NR_B200_(DEBUG_KNOBS|TUNING)", s)))
            elif re.match(r"#\s*endif", s) and depth:
                depth.pop()
            elif "getenv" in s and not s.startswith("//"):
                assert any(depth), "%s: getenv outside a knob guard: %s" % (name, s)
                guarded += 1
    from neural_renderer_b200 import build
    assert not any("NR_B200_DEBUG_KNOBS" in f or "NR_B200_TUNING" in f for f in build.NVCC_FLAGS)
