#!/bin/bash
# Emulates one rank of an N-GPU job on one GPU (tile_count=N, rank 0) and sweeps the host-side knobs.
V=";RT_SAMPLE_BUDGET_MB=3072;RT_SAMPLE_BUDGET_MB=8000;RT_PIPES_N=2"
for N in ${TILES:-8 4 2 1}; do RT_VARIANTS="$V" python tools/rank_time.py ${WL:-c4} $N 10 2>&1 | grep -v Warning; done
