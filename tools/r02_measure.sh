#!/bin/bash
# Round-2 measurement sequence on one B200 (under gpurun): GPU tests, smoke, both bench arms, the instruction /
# DRAM-byte counts bench.py's roofline uses, C5, every config next to the reference.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/final_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -9 gpurun_out/final_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; grep '^{' gpurun_out/final_bench_ref.json | cut -c1-160
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/final_bench_n1.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/final_bench_n1.json") if l.startswith("{")][0]
print(round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), d["frame_sha"][:12], "roofline", d["roofline"]["bound"], d["roofline"]["frac"], "cpu", round(d["cpu_baseline"]["value"],1), d["cpu_baseline"]["cores"])
for k,v in d["roofline"]["classes"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"], round(v.get("issue_frac_lane_weighted",0),3))
PY
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/final_inst.csv python bench.py --profile-frames 2 > gpurun_out/final_ncu_inst.log 2>&1; echo "inst rc=$?"
python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_c5.json 2> gpurun_out/final_bench_c5.err; echo "c5 rc=$?"
RT_DEVICE_BVH=1 python tools/all_configs.py > gpurun_out/final_all_configs.jsonl 2> gpurun_out/final_all_configs.err; grep '^{' gpurun_out/final_all_configs.jsonl | cut -c1-140
RT_TILES_LIST=1,2,4,8 RT_SLOTS=1,2,4 python tools/rank_overlap.py c4 16 2>&1 | grep "rank of" > gpurun_out/final_rank_overlap.txt; cat gpurun_out/final_rank_overlap.txt
