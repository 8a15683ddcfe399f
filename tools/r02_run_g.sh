#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/g_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/g_tests_gpu.log
for top in 0 1; do
RT_TOP_STAGE=$top python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_n1_top$top.json 2> gpurun_out/g_bench_n1_top$top.err; echo "bench top=$top rc=$?"; tail -2 gpurun_out/g_bench_n1_top$top.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/g_bench_n1_top$top.json") if l.startswith("{")][0]
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["frame_sha"][:12], "issue frac", d["roofline"]["frac"])
for k,v in d["roofline"]["classes"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"], round(v.get("issue_frac_lane_weighted",0),3))
PY
done
RT_TOP_STAGE=1 RT_SAMPLE_BUDGET_MB=3072 ncu --set full --clock-control none --import-source on -k regex:rt_walk_kernel -s 1 -c 1 -o gpurun_out/g_prof_walk_top python tools/render_once.py c4 4 1 > gpurun_out/g_ncu_walk_top.log 2>&1; echo "walk top rc=$?"
