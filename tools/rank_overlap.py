"""One rank's share of an N-GPU frame on ONE GPU (tile_count=N, rank 0): ms per frame when a frame starts after the
previous one has ended (one frame slot) and when consecutive frames alternate between the two frame slots
(rt_gpu_set_frame_slot: the next frame fills the wavefront's tail).   python tools/rank_overlap.py [workload] [frames]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import raytracerwin_b200 as rt
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 16
spec, W, H, passes, aa, bounce, mode, desc = bench.build_spec(wl)
scene = rt.Scene(spec)
pm = {"path": rt.RT_MODE_PATH, "preview": rt.RT_MODE_PREVIEW, "whitted": rt.RT_MODE_WHITTED}[mode]
if mode == "path": scene.set_unit_vectors(0, 0)
ctx = rt.GpuContext(0)
ctx.upload_scene(scene)
base = None
SLOTS = tuple(int(x) for x in os.environ.get("RT_SLOTS", "1,2,4").split(","))
for tiles in tuple(int(x) for x in os.environ.get("RT_TILES_LIST", "1,2,4,8").split(",")):
    tk = dict(tile_size=32, tile_count=tiles, tile_rank=0) if tiles > 1 else {}
    p = rt.make_params(W, H, mode=pm, max_bounce=bounce, pass_count=passes, antialias=aa, seed=0, **tk)
    res = {}
    for slots in SLOTS:
        def run(n):
            for k in range(n):
                if slots > 1: ctx.set_frame_slot(k % slots)
                ctx.reset_accum(W, H)
                ctx.render_tile(p)
            for s in range(slots):
                if slots > 1: ctx.set_frame_slot(s)
                ctx.synchronize()
        run(4)
        t0 = time.perf_counter(); run(frames); res[slots] = (time.perf_counter() - t0) / frames * 1e3
    if tiles == 1: base = res
    best1 = min(base.values()) if base else min(res.values()) * tiles
    print(f"{wl} rank of {tiles}: " + ", ".join(f"{n} slot(s) {res[n]:.3f} ms/frame (eff. vs best N=1 {best1 / tiles / res[n]:.3f})" for n in SLOTS), flush=True)
ctx.close()
