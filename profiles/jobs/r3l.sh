#!/bin/bash
# throughput vs batch size on one GPU (BASELINE fixes 4096 per GPU; this is the amortisation curve)
for b in 1024 2048 4096 8192 16384 32768; do
  timeout 600 python bench.py --batch $b --steps 10 --warmup 3 --no-extras --no-train --no-cpu-baseline > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err; echo "B=$b rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_b$b.json').read().strip().splitlines()[-1])
print('B=%6d: %.3f M q/s device-timed, %.3f ms per step, %d launches, e2e %.3f M q/s, phases %s' % ($b, d['value'] / 1e6, d['ms_per_step'], d['launches_per_step'], d['e2e']['value'] / 1e6, {k: round(v, 3) for k, v in (d.get('phases_ms') or {}).items()}))
PY
done 2>&1 | tee gpurun_out/batch_curve.txt
