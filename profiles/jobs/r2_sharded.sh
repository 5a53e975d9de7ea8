#!/bin/bash
# multi-GPU equality tests on hardware (run with gpurun --gpus N): sharded == unsharded, DP gradients == single process; short bench at N
N=$(nvidia-smi -L | wc -l)
nvidia-smi -L
python -m pytest tests/test_sharded_gpu.py -m gpu -q -s -rs > gpurun_out/r2_sharded_${N}gpu.log 2>&1; echo "sharded rc=$?"; tail -14 gpurun_out/r2_sharded_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "bench$N rc=$?"; tail -3 gpurun_out/bench_r2_n$N.err
python profiles/jobs/summarize_bench.py gpurun_out/bench_r2_n$N.json | cut -c1-900
