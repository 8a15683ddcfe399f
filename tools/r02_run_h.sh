#!/bin/bash
mkdir -p gpurun_out
for wl in c5 c5s c4; do
python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_$wl.json 2> gpurun_out/h_bench_$wl.err; echo "$wl bench rc=$?"; tail -2 gpurun_out/h_bench_$wl.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/h_bench_$wl.json") if l.startswith("{")][0]
print("$wl", round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["frame_sha"][:12], d["roofline"]["bound"], d["roofline"]["frac"])
for k,v in d["roofline"]["classes"].items(): print("   ",k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
done
timeout 900 python -m pytest tests/test_configs_at_size.py tests/test_c5_generated.py -m gpu -q -x > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/h_tests.log
