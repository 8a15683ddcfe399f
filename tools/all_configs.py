"""Every BASELINE config on this box: GPU (through the C-ABI, device-timed) next to the reference's CPU code on a
bounded sample.  Writes one JSON object per config to stdout.   python tools/all_configs.py [configs...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import raytracerwin_b200 as rt
import bench
from oracle import bindings

names = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5s", "c5"]
cores = os.cpu_count() or 1
ref = bindings.RefOracle() if bindings.ref_available() else None
ctx = rt.GpuContext(0)
for wl in names:
    spec, W, H, passes, aa, bounce, mode, desc = bench.build_spec(wl)
    pm = {"path": rt.RT_MODE_PATH, "preview": rt.RT_MODE_PREVIEW, "whitted": rt.RT_MODE_WHITTED}[mode]
    rt.use_device_bvh_builder(ctx if os.environ.get('RT_DEVICE_BVH') else None)
    t0 = time.perf_counter(); scene = rt.Scene(spec); t_load = time.perf_counter() - t0
    rt.use_device_bvh_builder(None)
    if mode == "path": scene.set_unit_vectors(0, 0)
    ctx.upload_scene(scene)
    p = rt.make_params(W, H, mode=pm, max_bounce=bounce, pass_count=passes, antialias=aa, seed=0)
    best = None
    for _ in range(4):
        ctx.reset_accum(W, H); ctx.reset_counters(); ctx.render_tile(p)
        ms = ctx.last_render_ms(); c = ctx.counters()
        best = ms if best is None else min(best, ms)
    out = {"config": wl, "desc": desc, "gpu_ms_per_frame": best, "rays_per_frame": c["rays"], "gpu_mrays_s": c["rays"] / best / 1e3,
           "camera_rays": c["camera_rays"], "visited_nodes_per_ray": c["node_visits"] / c["rays"], "triangles": scene.mesh_counts(len(spec) - 1)[3],
           "host_load_and_bvh_s": t_load}
    if ref is not None:
        # bounded CPU sample: one pass; for the generated scene only a band of rows (its rays are very long)
        if mode == "path": ref.init_unit_vectors(0)
        t0 = time.perf_counter(); rs = ref.build_scene(spec); t_ref_load = time.perf_counter() - t0
        rmode = {"path": 0, "preview": 1, "whitted": 2}[mode]
        start, end = (0, W * H - 1) if wl not in ("c5s", "c5") else ((H // 2 - 16) * W, (H // 2 + 16) * W - 1) if wl == "c5s" else ((H // 4) * W, (H // 4 + 16) * W - 1)
        r = ref.render(rs, W, H, mode=rmode, max_bounce=bounce, pass_begin=0, pass_count=1, antialias=aa, seed=0, nthreads=cores, start=start, end=end)
        # rays of exactly that sample from the device (same seed, same paths)
        ps = rt.make_params(W, H, mode=pm, max_bounce=bounce, pass_count=1, antialias=aa, seed=0, start=start, end=end)
        ctx.reset_accum(W, H); ctx.reset_counters(); ctx.render_tile(ps); cs = ctx.counters()
        same = bool(np.array_equal(np.nan_to_num(ctx.readback(rt.RT_READ_ACCUM_RGBN_F32, W, H)).view(np.uint32), np.nan_to_num(r["accum"]).view(np.uint32)))
        out.update({"cpu_sample": f"one pass, pixels {start}..{end}", "cpu_seconds": r["seconds"], "cpu_cores": cores,
                    "cpu_mrays_s": cs["rays"] / r["seconds"] / 1e6, "cpu_load_and_bvh_s": t_ref_load, "sample_bit_identical_to_reference": same,
                    "speedup": (c["rays"] / best / 1e3) / (cs["rays"] / r["seconds"] / 1e6)})
        ref.free_scene(rs)
    print(json.dumps(out), flush=True)
ctx.close()
