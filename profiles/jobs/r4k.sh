#!/bin/bash
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-train --no-cpu-baseline > gpurun_out/r3_bench_n2_async.json 2> gpurun_out/r3_bench_n2_async.err; echo "bench2 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3_bench_n2_async.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('parity'), d['e2e']['value'])
P
tail -3 gpurun_out/r3_bench_n2_async.err
