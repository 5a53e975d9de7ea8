#!/bin/bash
timeout 300 python profiles/micro_text_sort.py 2>&1 | tail -8 | tee gpurun_out/r3_micro_text_sort_v2.txt
B=4096 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_text_sort_launches.csv python profiles/micro_text_sort.py > gpurun_out/r3_ncu.log 2>&1
tail -3 gpurun_out/r3_ncu.log
