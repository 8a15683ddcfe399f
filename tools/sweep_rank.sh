#!/bin/bash
# Emulates one rank of an N-GPU job on one GPU (tile_count=N, rank 0) and sweeps the host-side knobs.
V=";RT_PACKET_ROUNDS=2;RT_PACKET_ROUNDS=2 RT_PACKET_MIN_LANES=16;RT_PACKET_ROUNDS=0"
for N in ${TILES:-1}; do RT_VARIANTS="$V" python tools/rank_time.py ${WL:-c4} $N 8 2>&1 | grep -v Warning; done
