import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import raytracerwin_b200 as rt, scenes
from oracle.bindings import PortOracle
D = os.path.join(ROOT, "assets/_ref/Data")
sc = rt.Scene(scenes.c3_unitychan(D))
W, H = 960, 540
ctx = rt.GpuContext(0); ctx.upload_scene(sc)
port = PortOracle()
for mode, aa in ((rt.RT_MODE_PREVIEW, 1), (rt.RT_MODE_PREVIEW, 0)):
  for tune in ((64, 20, 8, 8), (32, 1, 0, 1), (64, 32, 16, 32)):
    ctx.set_tuning(*tune)
    p = rt.make_params(W, H, mode=mode, antialias=aa, pass_count=1, seed=3)
    ctx.reset_accum(W, H); ctx.render_tile(p)
    g = ctx.readback(rt.RT_READ_ACCUM_RGBN_F32, W, H).copy()
    ctx.reset_accum(W, H); ctx.render_tile(p)
    g2 = ctx.readback(rt.RT_READ_ACCUM_RGBN_F32, W, H).copy()
    o = port.render(sc.desc, p, nthreads=16)["accum"]
    bad = (np.abs(g - o) > 1e-4).any(-1)
    ys, xs = np.nonzero(bad)
    print("mode", mode, "aa", aa, "tune", tune, "bad pixels", bad.sum(), "repeat-equal", np.array_equal(g.view(np.uint32), g2.view(np.uint32)))
    for y, x in list(zip(ys, xs))[:6]:
        print("  ", x, y, g[y, x], o[y, x])
