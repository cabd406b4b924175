#!/usr/bin/env python
"""Recipe that compiles the REFERENCE's own CUDA-C kernel strings into oracle/_ref/*.so.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package (neural_renderer_b200); only tests/, __graft_entry__.smoke()/build() and
bench.py's baseline legs use it, and only as the checker / the timed baseline.

What it does
------------
The reference (hiroharu-kato/neural_renderer) has no native sources: its device
code is seven CUDA-C strings inside ``neural_renderer/rasterize.py`` that CuPy
JIT-compiles (``chainer.cuda.elementwise`` -> ``cupy.ElementwiseKernel``).  CuPy and
Chainer are not installable here, so this script re-hosts those strings without
them, *reading them from where they lie under /root/reference at build time*:

  1. ``ast``-parse ``/root/reference/neural_renderer/rasterize.py`` and collect every
     ``chainer.cuda.elementwise(in_params, out_params, string.Template(BODY)
     .substitute(**kw), name)`` call site (rasterize.py:108-236 unsafe K3,
     :242-277 K1, :281-359 K2, :372-438 K4, :528-748 K5, :760-792 K6, :805-847 K7);
  2. substitute the same template keys the reference substitutes, with the same
     ``str()`` formatting (numeric constants are pasted as C literals, so
     ``near`` = ``0.1`` is a *double* literal and ``far`` = ``100`` an *int*);
  3. wrap each body exactly the way ``cupy.ElementwiseKernel`` does: a 1-D
     grid-stride loop ``for (ptrdiff_t i = tid; i < n; i += stride) { BODY }`` with
     ``raw`` arrays as plain pointers, launched with 128-thread blocks and
     ``ceil(n / 128)`` blocks;
  4. compile with ``nvcc -arch=sm_100a`` default flags (``--fmad=true``, no
     fast-math: what CuPy's NVRTC invocation uses) into
     ``oracle/_ref/nrref_<key>.so`` -- one shared object per configuration,
     because the reference bakes image_size / num_faces / near / far / eps /
     texture_size / return flags into the source.

Generated ``.cu`` text lives only in a temporary directory; no reference source
is copied into the repository.  ``oracle/_ref/`` is git-ignored (binaries travel
to the GPU box with the gpurun snapshot).  ``/root/reference`` does not exist on
the GPU box, so this script is run HERE (``__graft_entry__.build()`` calls it
when the reference tree is present).
"""
from __future__ import annotations

import ast
import hashlib
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("NR_REFERENCE_ROOT", "/root/reference")
REF_FILE = os.path.join(REF_ROOT, "neural_renderer", "rasterize.py")
OUT_DIR = os.path.join(HERE, "_ref")

# order of the elementwise call sites in rasterize.py (by line number)
KERNEL_NAMES = ["k3_unsafe", "k1_face_inv", "k2_zbuffer", "k4_texture", "k5_pixel_bwd", "k6_texture_bwd",
                "k7_depth_bwd"]

CTYPES = {"int32": "int", "float32": "float", "T": "float"}


def _const_str(node):
    """Evaluate a string expression made of constants and '+'."""
    if isinstance(node, ast.Constant) and isinstance(node.value, str):
        return node.value
    if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Add):
        return _const_str(node.left) + _const_str(node.right)
    raise ValueError("not a constant string expression: %s" % ast.dump(node))


def _is_elementwise(call):
    f = call.func
    return (isinstance(f, ast.Attribute) and f.attr == "elementwise"
            and isinstance(f.value, ast.Attribute) and f.value.attr == "cuda")


def collect_call_sites(path=REF_FILE, names=None):
    """Return [{name, in_params, out_params, template, keys, lineno}] in source order."""
    names = KERNEL_NAMES if names is None else names
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    sites = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and _is_elementwise(node):
            in_params = _const_str(node.args[0])
            out_params = _const_str(node.args[1])
            op = node.args[2]
            # string.Template('''...''').substitute(k=v, ...)
            assert isinstance(op, ast.Call) and op.func.attr == "substitute", ast.dump(op)[:200]
            tmpl_call = op.func.value
            assert isinstance(tmpl_call, ast.Call) and tmpl_call.func.attr == "Template"
            template = _const_str(tmpl_call.args[0])
            keys = [kw.arg for kw in op.keywords]
            sites.append(dict(in_params=in_params, out_params=out_params, template=template, keys=keys,
                              lineno=node.lineno))
    sites.sort(key=lambda s: s["lineno"])
    assert len(sites) == len(names), "reference layout changed: %d call sites" % len(sites)
    for s, n in zip(sites, names):
        s["name"] = n
    return sites


def _params(decl):
    """'int32 _, raw float32 faces' -> [('float', 'faces'), ...] (the non-raw loop placeholder is dropped)."""
    out = []
    for p in [x.strip() for x in decl.split(",") if x.strip()]:
        toks = p.split()
        if toks[0] != "raw":
            continue  # the 'int32 _' arange placeholder that only carries the element count
        out.append((CTYPES[toks[1]], toks[2]))
    return out


def config_key(cfg):
    s = json.dumps(cfg, sort_keys=True)
    tag = "S%d_F%d_ts%d_r%da%dd%d" % (cfg["image_size"], cfg["num_faces"], cfg["texture_size"],
                                      cfg["return_rgb"], cfg["return_alpha"], cfg["return_depth"])
    return tag + "_" + hashlib.sha1(s.encode()).hexdigest()[:8]


def normalize_config(image_size, num_faces, texture_size=0, near=0.1, far=100, eps=1e-4,
                     return_rgb=0, return_alpha=0, return_depth=0):
    # near/far/eps keep their Python type: the reference pastes str(value) into the source
    return dict(image_size=int(image_size), num_faces=int(num_faces), texture_size=int(texture_size),
                near=near, far=far, eps=eps, return_rgb=int(bool(return_rgb)),
                return_alpha=int(bool(return_alpha)), return_depth=int(bool(return_depth)))


def generate_source(cfg, sites):
    import string
    subst_all = dict(cfg)
    lines = ["#include <cuda_runtime.h>", "#include <stddef.h>", ""]
    for s in sites:
        if s["name"] == "k4_texture" and not cfg["texture_size"]:
            continue
        if s["name"] == "k6_texture_bwd" and not cfg["texture_size"]:
            continue
        kw = {k: subst_all[k] for k in s["keys"]}
        body = string.Template(s["template"]).substitute(**kw)
        params = _params(s["in_params"]) + _params(s["out_params"])
        sig = ", ".join("%s* %s" % (t, n) for t, n in params)
        lines.append('extern "C" __global__ void %s(%s, long long _n) {' % (s["name"], sig))
        lines.append("  for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < _n; "
                     "i += (ptrdiff_t)blockDim.x * gridDim.x) {")
        lines.append(body)
        lines.append("  }")
        lines.append("}")
        hsig = ", ".join("void* %s" % n for _, n in params)
        hargs = ", ".join("(%s*)%s" % (t, n) for t, n in params)
        lines.append('extern "C" int launch_%s(%s, long long n, void* stream) {' % (s["name"], hsig))
        lines.append("  if (n <= 0) return 0;")
        lines.append("  unsigned grid = (unsigned)((n + 127) / 128);")
        lines.append("  %s<<<grid, 128, 0, (cudaStream_t)stream>>>(%s, n);" % (s["name"], hargs))
        lines.append("  return (int)cudaGetLastError();")
        lines.append("}")
        lines.append("")
    return "\n".join(lines)


def build_one(cfg, sites=None, force=False, keep_ptx=None):
    sites = sites or collect_call_sites()
    os.makedirs(OUT_DIR, exist_ok=True)
    key = config_key(cfg)
    out = os.path.join(OUT_DIR, "nrref_%s.so" % key)
    if os.path.exists(out) and not force and not keep_ptx:
        return out
    src = generate_source(cfg, sites)
    with tempfile.TemporaryDirectory(prefix="nrref_") as td:
        cu = os.path.join(td, "ref.cu")
        with open(cu, "w") as f:
            f.write(src)
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared",
               "-Xcompiler", "-fPIC", "-w", "-o", out, cu]
        subprocess.run(cmd, check=True)
        if keep_ptx:
            subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-w", "-ptx",
                            "-o", keep_ptx, cu], check=True)
    with open(out + ".json", "w") as f:
        json.dump(cfg, f, sort_keys=True)
    return out


# ---- texture baking kernel of load_obj.py: one binary per (ts, image height, width)
REF_LOAD_OBJ = os.path.join(REF_ROOT, "neural_renderer", "load_obj.py")
BAKE_CONFIGS = [(4, 64, 48), (2, 64, 48), (6, 33, 57)]  # (texture_size, image_height, image_width) used by the tests


def bake_lib_path(texture_size, image_height, image_width):
    return os.path.join(OUT_DIR, "nrref_bake_ts%d_%dx%d.so" % (texture_size, image_height, image_width))


def build_bake(texture_size, image_height, image_width, force=False, keep_ptx=None):
    """The reference's bilinear texture-bake kernel string (load_obj.py:88-137), wrapped like the others."""
    out = bake_lib_path(texture_size, image_height, image_width)
    if os.path.exists(out) and not force and not keep_ptx:
        return out
    os.makedirs(OUT_DIR, exist_ok=True)
    sites = collect_call_sites(REF_LOAD_OBJ, names=["k_bake"])
    cfg = dict(image_size=0, num_faces=0, texture_size=int(texture_size), image_height=int(image_height),
               image_width=int(image_width))
    src = generate_source(cfg, sites)
    with tempfile.TemporaryDirectory(prefix="nrref_") as td:
        cu = os.path.join(td, "bake.cu")
        with open(cu, "w") as f:
            f.write(src)
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared",
                        "-Xcompiler", "-fPIC", "-w", "-o", out, cu], check=True)
        if keep_ptx:
            subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-w", "-cubin", "-o", keep_ptx, cu],
                           check=True)
    return out


def build_all(configs, force=False, jobs=None):
    if not os.path.exists(REF_FILE):
        raise FileNotFoundError(REF_FILE)
    sites = collect_call_sites()
    jobs = jobs or max(1, (os.cpu_count() or 2))
    with ThreadPoolExecutor(jobs) as ex:
        outs = list(ex.map(lambda c: build_one(c, sites, force), configs))
    return outs


def main(argv):
    sys.path.insert(0, HERE)
    from ref_configs import all_configs
    force = "--force" in argv
    cfgs = all_configs()
    outs = build_all(cfgs, force=force)
    if os.path.exists(REF_LOAD_OBJ):
        outs += [build_bake(*c, force=force) for c in BAKE_CONFIGS]
    print("built %d reference kernel libraries under %s" % (len(outs), OUT_DIR))


if __name__ == "__main__":
    main(sys.argv[1:])
