#!/bin/bash
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k "random_layout" 2>&1 | tail -30
