#!/bin/bash
N=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"lstm_fused|text_s|stage_rows_kernel<float|pair_kernel<5" --csv --log-file gpurun_out/r3_text_sort_launches_warm.csv python profiles/micro_text_sort.py > gpurun_out/r3_ncu.log 2>&1
tail -3 gpurun_out/r3_ncu.log
