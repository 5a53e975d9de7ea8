#!/bin/bash
for p in 0 1 0 1; do STAIR_LANE_PRIO=$p timeout 200 python profiles/micro_fwd_simple.py 2>&1 | tail -1; done | tee gpurun_out/r3_lane_prio_ab.txt
STAIR_LANE_PRIO=1 timeout 200 python profiles/module_timeline.py 4096 full 2>&1 | tail -32 > gpurun_out/r3_timeline_prio1.txt
for p in 0 1; do STAIR_LANE_PRIO=$p timeout 200 python profiles/micro_fwd_simple.py 4096 i3d 2>&1 | tail -1; done | tee -a gpurun_out/r3_lane_prio_ab.txt
