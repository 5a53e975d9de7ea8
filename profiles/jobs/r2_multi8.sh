#!/bin/bash
# 8-GPU call: multi-rank equality tests (2 / 4 / 8 ranks), H2D scaling diagnosis at N = 2, 4, 8, bench at N = 8
export STAIR_NGPU=8
nvidia-smi -L | head -8
python -m pytest tests/test_sharded_gpu.py -m gpu -q -s > gpurun_out/r2_sharded_8gpu.log 2>&1; echo "sharded rc=$?"; tail -12 gpurun_out/r2_sharded_8gpu.log
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) profiles/h2d_diag.py > gpurun_out/h2d_diag_n$n.txt 2>gpurun_out/h2d_diag_n$n.err; echo "h2d $n rc=$?"; head -6 gpurun_out/h2d_diag_n$n.txt
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; echo "bench8 rc=$?"
python profiles/jobs/summarize_bench.py gpurun_out/bench_r2_n8.json | cut -c1-700
