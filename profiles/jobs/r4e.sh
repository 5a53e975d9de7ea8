#!/bin/bash
# 2-GPU call on the final build: multi-rank equality tests, bench at N = 2
export STAIR_NGPU=2
nvidia-smi -L | head -2
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q -s > gpurun_out/r3_sharded_2gpu.log 2>&1; echo "sharded rc=$?"; tail -6 gpurun_out/r3_sharded_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3_bench_n2.json 2> gpurun_out/r3_bench_n2.err; echo "bench2 rc=$?"
python profiles/jobs/summarize_bench.py gpurun_out/r3_bench_n2.json | cut -c1-900
