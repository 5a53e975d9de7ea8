#!/bin/bash
TEXT_SORT=0 timeout 200 python profiles/module_timeline.py 4096 full 2>&1 | tail -32 > gpurun_out/r3_timeline_sort0.txt
TEXT_SORT=1 timeout 200 python profiles/module_timeline.py 4096 full 2>&1 | tail -32 > gpurun_out/r3_timeline_sort1.txt
