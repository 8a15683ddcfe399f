#!/bin/bash
# one-at-a-time sweep of the scheduling knobs (environment, read at context creation): C4 one frame at a time, and a rank's
# eighth with four frames in flight
run() { echo -n "$* : "; env "$@" python tools/render_once.py c4 16 4 2>/dev/null | grep "^c4" | awk '{print $11}' | sort -n | head -1 | tr '\n' ' '; env "$@" RT_TILES_LIST=8 RT_SLOTS=4 python tools/rank_overlap.py c4 16 2>/dev/null | grep "rank of" | awk '{print $7}'; }
run RT_NONE=0
for v in 16 32 48; do run RT_PACKET_PROBE=$v; done
for v in 6 8 12 14; do run RT_PACKET_MIN_LANES=$v; done
for v in 1024 4096; do run RT_LONG_LIMIT=$v; done
for v in 100000 400000; do run RT_THIN_COUNT=$v; done
for v in 128 512; do run RT_THIN_LIMIT=$v; done
for v in 12000 48000; do run RT_SMALL_ROUND=$v; done
for v in 300000 1200000; do run RT_THIN_GRID_COUNT=$v; done
# lane-per-walk kernel: refill threshold and leaf wait (rt_gpu_set_tuning through RT_TUNE = window,min_lanes,leaf_wait,pool_kpaths)
for ml in 24 28 31; do for lw in 6 12 18 24; do run RT_TUNE=32,$ml,$lw,0; done; done
