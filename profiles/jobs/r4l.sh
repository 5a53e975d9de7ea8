#!/bin/bash
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-train --no-cpu-baseline > gpurun_out/r3_bench_n8_async.json 2> gpurun_out/r3_bench_n8_async.err; echo "bench8 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3_bench_n8_async.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
P
