#!/bin/bash
# Emulates one rank of an N-GPU job on one GPU (tile_count=N, rank 0) and sweeps the host-side knobs.
V="RT_SMALL_ROUND=0;RT_SMALL_ROUND=24000"
for c in 200000 700000 2000000; do for l in 192 256 384; do V="$V;RT_SMALL_ROUND=24000 RT_THIN_COUNT=$c RT_THIN_LIMIT=$l"; done; done
for N in ${TILES:-8 1}; do RT_VARIANTS="$V" python tools/rank_time.py ${WL:-c4} $N 10 2>&1 | grep -v Warning; done
