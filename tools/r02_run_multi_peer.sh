#!/bin/bash
N=$1
RT_EXCHANGE=peer python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m_bench_peer_n$N.json 2> gpurun_out/m_bench_peer_n$N.err; echo "rc=$?"; tail -3 gpurun_out/m_bench_peer_n$N.err | cut -c1-200
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/m_bench_peer_n$N.json") if l.startswith("{")]
if d:
    d=d[0]; print("N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "sha", (d["frame_sha"] or "")[:12], d["run"]["exchange"][:40])
PY
