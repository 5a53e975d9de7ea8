#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/gemm_tests_r2l.log 2>&1; echo "gemm tests rc=$?"; tail -3 gpurun_out/gemm_tests_r2l.log
timeout 300 python profiles/micro_small_gemm.py > gpurun_out/micro_small_gemm.txt 2>&1; echo rc=$?; cat gpurun_out/micro_small_gemm.txt
timeout 300 python profiles/module_timeline.py 4096 full > gpurun_out/module_timeline_small.txt 2>&1; echo rc=$?; cat gpurun_out/module_timeline_small.txt
