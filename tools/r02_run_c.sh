#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/c_tests_gpu.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_n1.json 2> gpurun_out/c_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/c_bench_n1.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/c_bench_n1.json") if l.startswith("{")][0]
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["frame_sha"][:12])
for k,v in d["roofline"]["classes"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
RT_TILES_LIST=1,8 RT_SLOTS=1,4 python tools/rank_overlap.py c4 16 2>&1 | grep "rank of" > gpurun_out/c_rank_overlap.txt; cat gpurun_out/c_rank_overlap.txt
