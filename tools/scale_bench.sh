#!/bin/bash
# torchrun bench at N GPUs: tools/scale_bench.sh N tag [ENV=...]
n=$1; tag=$2; shift 2
env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n${n}_${tag}.json 2> gpurun_out/scale_n${n}_${tag}.err
grep '^{' gpurun_out/scale_n${n}_${tag}.json | cut -c1-260
