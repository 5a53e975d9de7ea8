#!/bin/bash
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k "length_sorted" 2>&1 | tail -8
