"""Renders c3 at reduced size with a tiny path pool (forces retry passes) and checks bit-equality with a big pool."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import raytracerwin_b200 as rt, scenes
D = os.path.join(ROOT, "assets/_ref/Data")
sc = rt.Scene(scenes.default_scene(D)); sc.set_unit_vectors(0, 0)
W, H = 640, 360
ctx = rt.GpuContext(0); ctx.upload_scene(sc)
outs = []
for pool_ki in (0, 64, 300):
    ctx.set_tuning(32, 28, 8, pool_ki)
    p = rt.make_params(W, H, mode=rt.RT_MODE_PATH, max_bounce=10, antialias=1, pass_count=2, seed=3)
    ctx.reset_accum(W, H); ctx.reset_counters(); ctx.render_tile(p)
    outs.append((ctx.readback(rt.RT_READ_ACCUM_RGBN_F32, W, H).copy(), ctx.counters()))
    print(pool_ki, ctx.last_render_ms(), outs[-1][1])
for o, c in outs[1:]:
    print("equal", np.array_equal(o.view(np.uint32), outs[0][0].view(np.uint32)), c == outs[0][1])
