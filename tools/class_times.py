"""Per-kernel-class time of one frame of a workload (single pipe, one event per launch): python tools/class_times.py c2"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import raytracerwin_b200 as rt
import bench
for wl in sys.argv[1:] or ["c1", "c2"]:
    spec, W, H, passes, aa, bounce, mode, desc = bench.build_spec(wl)
    scene = rt.Scene(spec)
    pm = {"path": rt.RT_MODE_PATH, "preview": rt.RT_MODE_PREVIEW, "whitted": rt.RT_MODE_WHITTED}[mode]
    if mode == "path": scene.set_unit_vectors(0, 0)
    ctx = rt.GpuContext(0)
    ctx.upload_scene(scene)
    p = rt.make_params(W, H, mode=pm, max_bounce=bounce, pass_count=passes, antialias=aa, seed=0)
    for i in range(3):
        ctx.reset_accum(W, H); ctx.render_tile(p)
    print(wl, "frame ms (4 pipes)", round(ctx.last_render_ms(), 4))
    ctx.set_pipes(1); ctx.time_kernels(2)
    for i in range(2):
        ctx.reset_accum(W, H); ctx.render_tile(p)
    print(wl, "frame ms (1 pipe, events)", round(ctx.last_render_ms(), 4), {k: (round(v[0], 4), v[1]) for k, v in ctx.kernel_class_ms().items()})
    ctx.close()
