#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/i_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/i_tests_gpu.log
for octo in 1 0; do
RT_OCTO=$octo python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_n1_octo$octo.json 2> gpurun_out/i_bench_n1_octo$octo.err; echo "bench octo=$octo rc=$?"; tail -2 gpurun_out/i_bench_n1_octo$octo.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/i_bench_n1_octo$octo.json") if l.startswith("{")][0]
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["frame_sha"][:12])
for k,v in d["roofline"]["classes"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
print("   visited nodes/ray", d["roofline"]["memory"]["visited_nodes_per_ray"])
PY
done
python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_c5.json 2> gpurun_out/i_bench_c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/i_bench_c5.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/i_bench_c5.json") if l.startswith("{")][0]
print("c5", round(d["value"]), round(d["ms_per_step"],2), d["frame_sha"][:12], d["run"]["scene_build_host_s"], d["run"]["scene_upload_s"])
for k,v in d["roofline"]["classes"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
