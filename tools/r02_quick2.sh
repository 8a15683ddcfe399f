#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/q_tests.log
for t in 1 0; do
RT_ORDERED_TREE=$t python bench.py --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/q_bench_tree$t.json 2> gpurun_out/q_bench_tree$t.err; echo "tree=$t rc=$?"; tail -1 gpurun_out/q_bench_tree$t.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/q_bench_tree$t.json") if l.startswith("{")][0]
print(round(d["value"]), round(d["ms_per_step"],2), d["frame_sha"][:12], "visited nodes/ray", round(d["roofline"]["memory"]["visited_nodes_per_ray"],2), "upload", round(d["run"]["scene_upload_s"],3))
for k,v in d["roofline"]["classes"].items(): print("   ",k, round(v["ms_per_step"],3))
PY
done
for t in 1 0; do RT_ORDERED_TREE=$t python tools/render_once.py c5s 4 3 2>/dev/null | tail -1; done
