#!/bin/bash
# quick check on one B200: GPU tests, one bench line with its class table, a rank's share of 8 with frames in flight
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/q_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q_bench_n1.json 2> gpurun_out/q_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/q_bench_n1.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/q_bench_n1.json") if l.startswith("{")][0]
print(round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), d["frame_sha"][:12], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step"]["frac"], "slots", d["run"]["frames_in_flight"])
for k,v in d["roofline"]["classes"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
RT_TILES_LIST=1,2,4,8 RT_SLOTS=1,4 python tools/rank_overlap.py c4 16 2>&1 | grep "rank of"
