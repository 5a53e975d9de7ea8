#!/bin/bash
timeout 300 python profiles/micro_interleave.py 2>&1 | tail -4 | tee gpurun_out/r3_micro_interleave.txt
timeout 200 python profiles/micro_collate.py 2>&1 | tail -22 | tee gpurun_out/r3_micro_collate.txt
