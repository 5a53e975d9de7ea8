#!/bin/bash
for pct in 100 97 92; do
STAIR_LANE_PRIO_PCT=$pct STAIR_LANE_PRIO=1 timeout 200 python profiles/micro_fwd_simple.py 2>&1 | tail -1
STAIR_LANE_PRIO_PCT=$pct STAIR_LANE_PRIO=1 timeout 200 python profiles/module_timeline.py 4096 full 2>&1 | tail -32 > gpurun_out/r3_timeline_prio_$pct.txt
done | tee gpurun_out/r3_lane_prio_ab2.txt
STAIR_LANE_PRIO=0 timeout 200 python profiles/micro_fwd_simple.py 2>&1 | tail -1 | tee -a gpurun_out/r3_lane_prio_ab2.txt
