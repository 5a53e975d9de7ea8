#!/bin/bash
# multi-GPU equality tests on hardware (run with gpurun --gpus N): sharded == unsharded, DP gradients == single process
nvidia-smi -L
python -m pytest tests/test_sharded_gpu.py -m gpu -q -s 2>&1 | tee gpurun_out/r2_sharded_${STAIR_NGPU:-N}gpu.log | tail -30
