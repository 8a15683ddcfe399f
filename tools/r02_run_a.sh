#!/bin/bash
# round 2, GPU call A: tests, smoke, first bench lines, instruction counts, shade / generate captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/a_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/a_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; tail -1 gpurun_out/a_smoke.log
python bench.py --steps 6 --warmup 3 > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err; echo "bench rc=$?"; grep '^{' gpurun_out/a_bench_n1.json | cut -c1-300; tail -3 gpurun_out/a_bench_n1.err
python bench.py --steps 6 --warmup 3 --no-overlap --no-cpu-baseline > gpurun_out/a_bench_n1_no_overlap.json 2> gpurun_out/a_bench_n1_no_overlap.err; grep '^{' gpurun_out/a_bench_n1_no_overlap.json | cut -c1-200
python tools/rank_overlap.py c4 16 > gpurun_out/a_rank_overlap.txt 2>&1; cat gpurun_out/a_rank_overlap.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; grep '^{' gpurun_out/a_bench_ref.json | cut -c1-200
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/a_inst.csv python bench.py --profile-frames 2 > gpurun_out/a_ncu_inst.log 2>&1; echo "inst rc=$?"
RT_SAMPLE_BUDGET_MB=3072 ncu --set full --clock-control none --import-source on -k regex:rt_shade_kernel -s 1 -c 1 -o gpurun_out/a_prof_shade python tools/render_once.py c4 4 1 > gpurun_out/a_ncu_shade.log 2>&1; echo "shade rc=$?"
RT_SAMPLE_BUDGET_MB=3072 ncu --set full --clock-control none --import-source on -k regex:rt_generate_kernel -c 1 -o gpurun_out/a_prof_generate python tools/render_once.py c4 4 1 > gpurun_out/a_ncu_gen.log 2>&1; echo "gen rc=$?"
