#!/bin/bash
for cfg in "4 0" "8 0" "4 1" "6 1" "8 1"; do set -- $cfg
  BWD_LANES=$1 DEP=$2 timeout 200 python profiles/micro_train_phases.py 2>&1 | head -4
done > gpurun_out/train_bwd_sched.txt 2>&1
cat gpurun_out/train_bwd_sched.txt
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -3
