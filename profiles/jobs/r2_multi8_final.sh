#!/bin/bash
# 8-GPU call, final build: multi-rank equality tests (2 / 4 / 8 ranks), bench at N = 2, 4, 8
export STAIR_NGPU=8
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -q -s > gpurun_out/r2_sharded_8gpu_final.log 2>&1; echo "sharded rc=$?"; tail -8 gpurun_out/r2_sharded_8gpu_final.log
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_final_n$n.json 2> gpurun_out/bench_final_n$n.err; echo "bench$n rc=$?"
  python profiles/jobs/summarize_bench.py gpurun_out/bench_final_n$n.json | grep "^value\|^ms_per_step\|^e2e\|^train\|^parity" | cut -c1-420
done
