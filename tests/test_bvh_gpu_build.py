"""KdTree::Build on the device == the host builder (itself pinned to the reference's tree by
tests/test_oracle_vs_ref.py), bit for bit: node bounds, escape links, leaf slots, triangle records, depth."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_c5  # noqa: E402

pytestmark = pytest.mark.gpu


def check(rt, gpu, sc):
    d = sc.mesh_dump(0)
    nodes, tris, shade = sc.flat_mesh(0)
    gn, gt, depth, ms = gpu.build_bvh(d["points"], d["pidx"])
    assert depth == sc.mesh_counts(0)[6]
    assert gn.tobytes() == nodes.tobytes()
    assert gt.tobytes() == tris.tobytes()
    return ms


@pytest.mark.parametrize("name", ["TorusKnot", "BlenderMonkey", "unitychan"])
def test_gpu_build_equals_host_build(rt, gpu, data_dir, name):
    sc = rt.Scene([("mesh", f"{data_dir}/{name}.obj", ("diffuse", (1, 1, 1)))])
    check(rt, gpu, sc)


def test_gpu_build_generated_grid_and_degenerates(tmp_path, rt, gpu, data_dir):
    out = str(tmp_path / "torus_200.obj")
    make_c5.main(os.path.join(data_dir, "TorusKnot.obj"), out, 200)           # 240 k triangles, "-0.000000" coordinates included
    sc = rt.Scene([("mesh", out, ("diffuse", (1, 1, 1)))])
    ms = check(rt, gpu, sc)
    assert ms > 0
    # coincident centroids (half-split fallback), zero-area and repeated triangles, signed zeros
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [-0.0, 0.0, -0.0], [0.0, -0.0, 0.0], [2, 2, 2], [2, 2, 2], [2, 2, 2],
                    [1, 1, -0.0], [-1, 0.0, 0.0], [0.0, -1, 0.0]], np.float32)
    idx = np.array([[0, 1, 2], [0, 1, 2], [0, 1, 2], [3, 4, 3], [5, 6, 7], [0, 1, 2], [8, 9, 10], [3, 1, 2], [4, 1, 2], [0, 1, 2]], np.int32)
    sc2 = rt.Scene()
    sc2.add_mesh_arrays(pts, idx, material=("diffuse", (1, 1, 1)))
    check(rt, gpu, sc2)
    sc3 = rt.Scene()
    sc3.add_mesh_arrays(pts, idx[:1], material=("diffuse", (1, 1, 1)))      # a single triangle: the root is a leaf
    check(rt, gpu, sc3)


def test_scene_loader_with_device_builder(rt, gpu, port, data_dir):
    """The host scene layer with the device builder switched on: same flattened mesh, same render."""
    spec = [("mesh", f"{data_dir}/unitychan.obj", ("reflective", (0.8, 0.8, 0.8), 0.0))]
    host = rt.Scene(spec)
    rt.use_device_bvh_builder(gpu)
    try:
        dev = rt.Scene(spec)
    finally:
        rt.use_device_bvh_builder(None)
    for a, b in zip(host.flat_mesh(0), dev.flat_mesh(0)):
        assert a.tobytes() == b.tobytes()
    assert host.mesh_counts(0) == dev.mesh_counts(0)
    gpu.upload_scene(dev)
    W, H = 320, 180
    p = rt.make_params(W, H, mode=rt.RT_MODE_PATH, max_bounce=4, antialias=0)
    gpu.reset_accum(W, H)
    gpu.render_tile(p)
    acc = gpu.readback(rt.RT_READ_ACCUM_RGBN_F32, W, H)
    o = port.render(host.desc, p, nthreads=8)
    assert acc.tobytes() == o["accum"].tobytes()
