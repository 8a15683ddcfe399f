#!/bin/bash
# C5 (10 M triangles): bench line with the HBM roofline, ncu capture of the walk kernel, CPU reference beside it
mkdir -p gpurun_out
python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_c5.json 2> gpurun_out/e_bench_c5.err; echo "c5 bench rc=$?"; tail -3 gpurun_out/e_bench_c5.err; grep '^{' gpurun_out/e_bench_c5.json | cut -c1-300
python tools/render_once.py c5 4 1 > gpurun_out/e_render_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rt_walk_kernel -s 1 -c 1 -o gpurun_out/e_prof_walk_c5 python tools/render_once.py c5 4 1 > gpurun_out/e_ncu_walk_c5.log 2>&1; echo "ncu walk c5 rc=$?"; tail -2 gpurun_out/e_render_c5.log
RT_DEVICE_BVH=1 python tools/all_configs.py c1 c2 c3 c4 c5s c5 > gpurun_out/e_all_configs.jsonl 2> gpurun_out/e_all_configs.err; grep '^{' gpurun_out/e_all_configs.jsonl | cut -c1-200; tail -2 gpurun_out/e_all_configs.err
