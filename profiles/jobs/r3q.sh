#!/bin/bash
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k "autograd" 2>&1 | tail -40
