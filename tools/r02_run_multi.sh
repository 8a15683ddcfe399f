#!/bin/bash
# usage: r02_run_multi.sh N [extra bench args]   (under gpurun --gpus N)
N=$1; shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err; echo "bench N=$N rc=$?"; tail -5 gpurun_out/m_bench_n$N.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/m_bench_n$N.json") if l.startswith("{")]
if d:
    d=d[0]; print("N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "sha", (d["frame_sha"] or "")[:12], "slots", d["run"]["frames_in_flight"], d["run"]["exchange"][:30])
PY
