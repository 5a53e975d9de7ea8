#!/bin/bash
# ncu --set full of the length-sorted recurrence kernel inside a steady-state forward (after the program ran clean without ncu)
timeout 200 python profiles/micro_fwd_simple.py > gpurun_out/r3_fwd_plain_v2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lstm_fused -s 6 -c 1 -o gpurun_out/r3_lstm_sorted_v2 python profiles/micro_fwd_simple.py > gpurun_out/r3_ncu_lstm_v2.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r3_lstm_sorted_v2.ncu-rep --page raw --csv > gpurun_out/r3_lstm_sorted_v2_raw.csv 2>/dev/null; ls -la gpurun_out/r3_lstm_sorted_v2*; cat gpurun_out/r3_fwd_plain_v2.log
