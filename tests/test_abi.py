"""The C-ABI library loads, exports every symbol the headers declare, and its struct layouts match the
ctypes mirror.  No compute calls: runs on a CPU-only box."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from conftest import ROOT

INCLUDE = os.path.join(ROOT, "include")


def declared_functions(header):
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_\s\*]*?\b(rt_(?:gpu|host)_[a-z0-9_]+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_headers_declare_what_the_bindings_bind(rt):
    abi = rt._abi
    gpu = declared_functions("rt_gpu.h")
    host = declared_functions("rt_host.h")
    assert len(gpu) >= 20 and len(host) >= 25
    assert set(abi.GPU_PROTOTYPES) <= set(gpu), set(abi.GPU_PROTOTYPES) - set(gpu)
    assert set(abi.HOST_PROTOTYPES) <= set(host), set(abi.HOST_PROTOTYPES) - set(host)


def test_library_exports_every_declared_symbol(rt):
    lib = C.CDLL(rt._abi.lib_path())
    for header in ("rt_gpu.h", "rt_host.h"):
        for name in declared_functions(header):
            assert hasattr(lib, name), f"{name} declared in {header} but not exported"
    assert lib.rt_gpu_abi_version() == rt._abi.RT_GPU_ABI_VERSION


def test_library_exports_nothing_the_headers_do_not_declare(rt):
    """The other direction: every dynamic symbol the library defines is declared in include/*.h (the tooling
    entry points live in rt_gpu_debug.h); kernels' host stubs and C++ internals stay local (csrc/exports.map)."""
    out = subprocess.run(["nm", "-D", "--defined-only", rt._abi.lib_path()], check=True, capture_output=True, text=True).stdout
    exported = sorted({line.split()[-1] for line in out.splitlines() if line.split() and line.split()[-2] in ("T", "t", "W", "B", "D")})
    declared = set()
    for header in ("rt_gpu.h", "rt_host.h", "rt_gpu_debug.h"):
        declared |= set(declared_functions(header))
    assert exported, "nm found no exported symbols"
    assert set(exported) <= declared, sorted(set(exported) - declared)


def test_struct_layouts_match_c(rt):
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "rt_gpu.h"
#include "rt_host.h"
int main(void) {
  printf("rt_shape %zu\n", sizeof(rt_shape)); printf("rt_material %zu\n", sizeof(rt_material));
  printf("rt_bvh_node %zu\n", sizeof(rt_bvh_node)); printf("rt_tri %zu\n", sizeof(rt_tri));
  printf("rt_shade %zu\n", sizeof(rt_shade)); printf("rt_texture %zu\n", sizeof(rt_texture));
  printf("rt_mesh %zu\n", sizeof(rt_mesh)); printf("rt_light %zu\n", sizeof(rt_light));
  printf("rt_scene_desc %zu\n", sizeof(rt_scene_desc)); printf("rt_render_params %zu\n", sizeof(rt_render_params));
  printf("rt_counters %zu\n", sizeof(rt_counters));
  printf("off_scene_eye %zu\n", offsetof(rt_scene_desc, eye)); printf("off_params_tile_rank %zu\n", offsetof(rt_render_params, tile_rank));
  printf("off_mesh_textures %zu\n", offsetof(rt_mesh, textures)); printf("off_shape_radius %zu\n", offsetof(rt_shape, radius));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "sizes.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "sizes")
        subprocess.run(["gcc", "-std=c11", "-I", INCLUDE, c, "-o", exe], check=True)     # the headers are plain C
        out = dict(line.split() for line in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.splitlines())
    abi = rt._abi
    for name in ("rt_shape", "rt_material", "rt_bvh_node", "rt_tri", "rt_shade", "rt_texture", "rt_mesh", "rt_light",
                 "rt_scene_desc", "rt_render_params", "rt_counters"):
        assert C.sizeof(getattr(abi, name)) == int(out[name]), name
    assert abi.rt_scene_desc.eye.offset == int(out["off_scene_eye"])
    assert abi.rt_render_params.tile_rank.offset == int(out["off_params_tile_rank"])
    assert abi.rt_mesh.textures.offset == int(out["off_mesh_textures"])
    assert abi.rt_shape.radius.offset == int(out["off_shape_radius"])
    for name, size in abi.STRUCT_SIZES.items():
        assert int(out[name]) == size


def test_no_cpu_fallback(rt):
    """Without a CUDA device the render path refuses to exist (there is nothing to fall back to)."""
    lib = rt.load_library()
    if lib.rt_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(rt.RtError) as e:
        rt.GpuContext(0)
    assert "no usable CUDA device" in str(e.value)
    h = C.c_void_p()
    assert lib.rt_gpu_create(0, C.byref(h)) == rt._abi.RT_ERR_CUDA and not h.value


def test_product_never_touches_the_checkers():
    """Only tests/, bench.py's baseline legs and __graft_entry__.smoke() may use oracle/: the product
    sources neither import, link nor load anything from there."""
    pkg = os.path.join(ROOT, "raytracerwin_b200")
    for dirpath, dirs, files in os.walk(pkg):
        dirs[:] = [d for d in dirs if d not in ("build", "__pycache__")]
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("librt_oracle", "libref_oracle", "oracle.bindings", "from oracle", "import oracle", "oracle/_ref", "-lrt_oracle"):
                    assert needle not in text, f"{os.path.join(dirpath, f)} references {needle}"


def test_owned_pixels_partition(rt):
    for (W, H, T, n) in ((3840, 2160, 32, 8), (1920, 1080, 32, 4), (150, 70, 32, 3), (33, 17, 8, 5), (64, 64, 64, 2)):
        counts = [rt.owned_pixels(W, H, T, n, r) for r in range(n)]
        assert sum(counts) == W * H
    counts = [rt.owned_pixels(3840, 2160, 32, 8, r) for r in range(8)]
    assert max(counts) / min(counts) < 1.01          # the interleave is balanced in area
    assert rt.owned_pixels(100, 50, 0, 0, 0) == 5000
