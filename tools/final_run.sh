#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/final_tests_gpu.log 2>&1; tail -2 gpurun_out/final_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; grep '^{' gpurun_out/final_bench_n1.json | cut -c1-200
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1; echo launches rc=$?
RT_SAMPLE_BUDGET_MB=3072 python tools/render_once.py c4 4 2 > /dev/null 2>&1 && RT_SAMPLE_BUDGET_MB=3072 ncu --set full --clock-control none --import-source on -k regex:rt_walk_kernel -s 41 -c 1 -o gpurun_out/final_prof_walk python tools/render_once.py c4 4 2 > /dev/null 2>&1; echo walk rc=$?
RT_SAMPLE_BUDGET_MB=3072 ncu --set full --clock-control none --import-source on -k regex:rt_walk_packet_kernel -s 4 -c 1 -o gpurun_out/final_prof_packet python tools/render_once.py c4 4 2 > /dev/null 2>&1; echo packet rc=$?
RT_DEVICE_BVH=1 python tools/all_configs.py > gpurun_out/final_all_configs.jsonl 2> gpurun_out/final_all_configs.err; grep '^{' gpurun_out/final_all_configs.jsonl | cut -c1-160
