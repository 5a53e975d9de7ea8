#!/bin/bash
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -8
STAIR_TEXT_SORT=0 timeout 300 python profiles/micro_train_phases.py 2>&1 | tail -8 | tee gpurun_out/r3_train_phases_sort0.txt
STAIR_TEXT_SORT=1 timeout 300 python profiles/micro_train_phases.py 2>&1 | tail -8 | tee gpurun_out/r3_train_phases_sort1.txt
