"""Small driver for profiling: render a workload a few times through the C-ABI and print timings.
   python tools/render_once.py [workload] [passes] [repeats] [exact]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import raytracerwin_b200 as rt
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
spec, W, H, passes, aa, bounce, mode, desc = bench.build_spec(wl)
if len(sys.argv) > 2: passes = int(sys.argv[2])
repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 3
trav = rt.RT_TRAVERSE_EXACT if len(sys.argv) > 4 and sys.argv[4] == "exact" else rt.RT_TRAVERSE_CULLED
scene = rt.Scene(spec)
pm = {"path": rt.RT_MODE_PATH, "preview": rt.RT_MODE_PREVIEW, "whitted": rt.RT_MODE_WHITTED}[mode]
if mode == "path": scene.set_unit_vectors(0, 0)
ctx = rt.GpuContext(0)
ctx.upload_scene(scene)
ctx.time_kernels(True)
if os.environ.get('RT_PIPES_N'): ctx.set_pipes(int(os.environ['RT_PIPES_N']))
if os.environ.get('RT_TUNE'):
    ctx.set_tuning(*map(int, os.environ['RT_TUNE'].split(',')))
tk = dict(tile_size=int(os.environ.get('RT_TILE_SIZE', '32')), tile_count=int(os.environ['RT_TILES']), tile_rank=int(os.environ.get('RT_RANK', '0'))) if os.environ.get('RT_TILES') else {}
p = rt.make_params(W, H, mode=pm, max_bounce=bounce, pass_count=passes, antialias=aa, seed=0, traverse=trav, **tk)
for i in range(repeats):
    ctx.reset_accum(W, H); ctx.reset_counters()
    ctx.render_tile(p)
    ms, n = ctx.last_kernel_ms()
    c = ctx.counters()
    print(f"{wl} {W}x{H} passes={passes} kernel {ms:.3f} ms in {n} launch(es), total {ctx.last_render_ms():.3f} ms, "
          f"{c['rays']/ms/1e3:.1f} Mrays/s, rays {c['rays']}, nodes/ray {c['node_visits']/c['rays']:.2f}, tris/ray {c['tri_visits']/c['rays']:.3f}")
if os.environ.get('RT_ROUNDS'):
    import ctypes as C
    n = int(os.environ['RT_ROUNDS'])
    cnt = (C.c_uint32 * n)(); ms = (C.c_float * n)()
    lib = rt.load_library(); lib.rt_gpu_debug_rounds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    k = lib.rt_gpu_debug_rounds(ctx.handle, cnt, ms, n)
    print('walk launches', k)
    for i in range(n - 1): print(f'  round {i}: entries {cnt[i]:9d}  walk {ms[i]:.3f} ms')
    print('  longest single walk (nodes):', cnt[n - 1])
    lc = (C.c_uint32 * n)(); lib.rt_gpu_debug_long.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]; lib.rt_gpu_debug_long(ctx.handle, lc, n)
    print('  long walks per round:', list(lc))
ctx.close()
