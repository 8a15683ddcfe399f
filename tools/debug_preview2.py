import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import raytracerwin_b200 as rt, scenes
D = os.path.join(ROOT, "assets/_ref/Data")
sc = rt.Scene(scenes.c3_unitychan(D))
W, H = 480, 270
ctx = rt.GpuContext(0); ctx.upload_scene(sc)
p = rt.make_params(W, H, mode=rt.RT_MODE_PREVIEW, antialias=0, pass_count=1, seed=3)
ctx.reset_accum(W, H); ctx.render_tile(p)
g = ctx.readback(rt.RT_READ_ACCUM_RGBN_F32, W, H).copy()
print("max", g[..., :3].max())
