#!/bin/bash
timeout 300 python profiles/micro_text_sort.py 2>&1 | tail -5 | tee gpurun_out/r3_micro_text_sort_v5.txt
