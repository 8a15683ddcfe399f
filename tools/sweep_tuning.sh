#!/bin/bash
# walk-kernel scheduling knobs on C4 (one frame at a time, best of 4): RT_TUNE = window,min_lanes,leaf_wait,pool_kpaths
for ml in 24 28 31; do for lw in 6 12 18 24; do
  echo -n "min_lanes=$ml leaf_wait=$lw: "
  RT_TUNE=32,$ml,$lw,0 python tools/render_once.py c4 16 4 2>/dev/null | grep "^c4" | awk '{print $11}' | sort -n | head -1
done; done
