#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/b_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/b_tests_gpu.log
for fr in 0 2 3 4; do echo "RT_FINISH_ROUND=$fr"; RT_FINISH_ROUND=$fr python tools/rank_overlap.py c4 16 2>&1 | grep -v "^Load\|^Mesh\|^Gener"; done > gpurun_out/b_finish_round.txt 2>&1; cat gpurun_out/b_finish_round.txt
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_n1.json 2> gpurun_out/b_bench_n1.err; echo "bench rc=$?"; grep '^{' gpurun_out/b_bench_n1.json | cut -c1-200; tail -3 gpurun_out/b_bench_n1.err
python - <<'PY'
import time, sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import raytracerwin_b200 as rt, scenes
t0 = time.perf_counter(); sc = rt.Scene(scenes.c3_unitychan("assets/_ref/Data")); print("unitychan load s", time.perf_counter() - t0)
t0 = time.perf_counter(); sc = rt.Scene(scenes.c3_unitychan("assets/_ref/Data")); print("unitychan load s (2nd)", time.perf_counter() - t0)
ctx = rt.GpuContext(0); t0 = time.perf_counter(); ctx.upload_scene(sc); ctx.synchronize(); print("upload s", time.perf_counter() - t0)
t0 = time.perf_counter(); ctx.upload_scene(sc); ctx.synchronize(); print("upload s (2nd)", time.perf_counter() - t0)
PY
for fr in 0 -1 1; do echo "RT_FINISH_ROUND=$fr"; RT_FINISH_ROUND=$fr python tools/all_configs.py c1 c2 c3 2>/dev/null | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'], 'gpu ms', round(d['gpu_ms_per_frame'], 4), 'Mrays/s', round(d['gpu_mrays_s']), 'cpu', round(d.get('cpu_mrays_s', 0), 1), 'x', round(d.get('speedup', 0), 1), 'same', d.get('sample_bit_identical_to_reference'))
"; done > gpurun_out/b_small_frames.txt 2>&1; cat gpurun_out/b_small_frames.txt
