#!/bin/bash
# 8-GPU call on the final build: multi-rank equality tests (2 / 4 / 8 ranks), bench at N = 8
export STAIR_NGPU=8
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q -s > gpurun_out/r3_sharded_8gpu.log 2>&1; echo "sharded rc=$?"; tail -4 gpurun_out/r3_sharded_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r3_bench_n8.json 2> gpurun_out/r3_bench_n8.err; echo "bench8 rc=$?"
python profiles/jobs/summarize_bench.py gpurun_out/r3_bench_n8.json | cut -c1-400 | head -8
