#!/bin/bash
timeout 600 python -m pytest tests/test_edge_gpu.py -m gpu -q -x -k "length_sorted or optional_kernel" 2>&1 | tail -15
timeout 300 python profiles/micro_text_sort.py 2>&1 | tail -8 | tee gpurun_out/r3_micro_text_sort.txt
