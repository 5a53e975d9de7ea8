#!/bin/bash
timeout 600 python -m pytest tests/test_edge_gpu.py -m gpu -q -x -k "optional_kernel_forms" 2>&1 | tail -12
