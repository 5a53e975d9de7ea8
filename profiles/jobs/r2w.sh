#!/bin/bash
python profiles/micro_i3d_fwd.py > gpurun_out/i3d_fwd_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1100 --csv --log-file gpurun_out/launches_i3d_warm.csv python profiles/micro_i3d_fwd.py > gpurun_out/ncu_i3d.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/i3d_fwd_plain.log
