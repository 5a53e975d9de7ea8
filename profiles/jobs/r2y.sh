#!/bin/bash
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_operators_gpu.py tests/test_edge_gpu.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -70
