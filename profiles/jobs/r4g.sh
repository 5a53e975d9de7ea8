#!/bin/bash
timeout 200 python profiles/micro_fwd_simple.py 2>&1 | tail -1
timeout 200 python profiles/micro_fwd_simple.py 4096 i3d 2>&1 | tail -1
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_edge_gpu.py tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -4
STAIR_TEXT_SORT=1 timeout 300 python profiles/micro_train_phases.py 2>&1 | tail -5
