"""Walk-bracket timeline of one call (which pipe's round ran when): python tools/timeline.py [workload] [tiles]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import raytracerwin_b200 as rt
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
spec, W, H, passes, aa, bounce, mode, desc = bench.build_spec(wl)
scene = rt.Scene(spec)
pm = {"path": rt.RT_MODE_PATH, "preview": rt.RT_MODE_PREVIEW, "whitted": rt.RT_MODE_WHITTED}[mode]
if mode == "path": scene.set_unit_vectors(0, 0)
ctx = rt.GpuContext(0)
ctx.upload_scene(scene)
ctx.time_kernels(True)
tk = dict(tile_size=32, tile_count=tiles, tile_rank=0) if tiles > 1 else {}
p = rt.make_params(W, H, mode=pm, max_bounce=bounce, pass_count=passes, antialias=aa, seed=0, **tk)
for i in range(4):
    ctx.reset_accum(W, H)
    ctx.render_tile(p)
    total = ctx.last_render_ms()
lib = rt.load_library()
cap = 4096
b = (C.c_float * cap)(); e = (C.c_float * cap)()
lib.rt_gpu_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
n = lib.rt_gpu_debug_timeline(ctx.handle, b, e, cap)
rounds = bounce if mode == "path" else 1
print(f"{wl} tiles={tiles} total {total:.3f} ms, {n} brackets")
for k in range(n):
    print(f"  chunk {k // rounds} round {k % rounds:2d}: {b[k]:8.3f} -> {e[k]:8.3f}  ({e[k]-b[k]:.3f})")
ctx.close()
