#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench_n1.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/f_bench_n1.json") if l.startswith("{")][0]
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["frame_sha"][:12])
for k,v in d["roofline"]["classes"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/f_inst.csv python bench.py --profile-frames 2 > gpurun_out/f_ncu_inst.log 2>&1; echo "inst rc=$?"
RT_SAMPLE_BUDGET_MB=3072 ncu --set full --clock-control none --import-source on -k regex:rt_walk_kernel -s 1 -c 1 -o gpurun_out/f_prof_walk python tools/render_once.py c4 4 1 > gpurun_out/f_ncu_walk.log 2>&1; echo "walk rc=$?"
RT_SAMPLE_BUDGET_MB=3072 ncu --set full --clock-control none --import-source on -k regex:rt_shade_kernel -s 1 -c 1 -o gpurun_out/f_prof_shade python tools/render_once.py c4 4 1 > gpurun_out/f_ncu_shade.log 2>&1; echo "shade rc=$?"
bash tools/r02_run_e.sh
