#!/bin/bash
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_edge_gpu.py -m gpu -q -x 2>&1 | tail -3
LANES=8 timeout 200 python profiles/micro_fwd.py 2>&1 | tail -2
LANES=8 timeout 200 python profiles/micro_fwd.py 2>&1 | tail -2
