"""4K frame with MaxBounceTimes 32: the path pools are capped by free device memory, retry passes take over."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import raytracerwin_b200 as rt, bench
spec, W, H, passes, aa, bounce, mode, desc = bench.build_spec("c4")
sc = rt.Scene(spec); sc.set_unit_vectors(0, 0)
ctx = rt.GpuContext(0); ctx.upload_scene(sc)
for b in (10, 32, 10):
    p = rt.make_params(W, H, mode=rt.RT_MODE_PATH, max_bounce=b, pass_count=4, antialias=1, seed=0)
    ctx.reset_accum(W, H); ctx.reset_counters(); ctx.render_tile(p)
    print("max_bounce", b, "ms", round(ctx.last_render_ms(), 2), "rays", ctx.counters()["rays"])
import torch
print("device memory in use (GB):", round((torch.cuda.mem_get_info(0)[1] - torch.cuda.mem_get_info(0)[0]) / 1e9, 1))
