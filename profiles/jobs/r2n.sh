#!/bin/bash
# warm-cache serialised launch list of one forward (ncu --cache-control none): per-kernel durations without the cold-cache penalty
export LANES=8
python profiles/micro_fwd.py > gpurun_out/fwd_plain_r2n.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1400 --csv --log-file gpurun_out/launches_fwd_warm.csv python profiles/micro_fwd.py > gpurun_out/ncu_fwd_warm.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/fwd_plain_r2n.log
