#!/bin/bash
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_loops_gpu.py tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/train_tests_r2r.log 2>&1; echo "train tests rc=$?"; tail -4 gpurun_out/train_tests_r2r.log
timeout 300 python profiles/micro_train_host.py 32 > gpurun_out/train_host32_v2.txt 2>&1; echo rc=$?; cat gpurun_out/train_host32_v2.txt
