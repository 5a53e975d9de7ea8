#!/bin/bash
timeout 900 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "random_layouts" 2>&1 | tail -30
