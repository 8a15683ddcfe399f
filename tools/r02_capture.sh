#!/bin/bash
# ncu --set full captures of single launches (round 1 of a 4-pass 4K chunk; the first generate launch)
mkdir -p gpurun_out
RT_SAMPLE_BUDGET_MB=3072 python tools/render_once.py c4 4 1 > /dev/null 2>&1
for k in generate walk shade; do
  skip=1; [ $k = generate ] && skip=0
  RT_SAMPLE_BUDGET_MB=3072 ncu --set full --clock-control none --import-source on -k regex:rt_${k}_kernel -s $skip -c 1 -o gpurun_out/cap_$k python tools/render_once.py c4 4 1 > gpurun_out/cap_$k.log 2>&1; echo "$k rc=$?"
done
