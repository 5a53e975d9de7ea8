#!/bin/bash
STAIR_TEXT_SORT=0 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r3_bench_sort0.json 2> gpurun_out/r3_bench_sort0.err
STAIR_TEXT_SORT=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r3_bench_sort1.json 2> gpurun_out/r3_bench_sort1.err
python - <<'P'
import json
for f in ('gpurun_out/r3_bench_sort0.json','gpurun_out/r3_bench_sort1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d.get('phases_ms'), d['e2e']['value'], d.get('parity',{}) if isinstance(d.get('parity'),dict) and len(str(d.get('parity')))<600 else '')
    except Exception as e:
        print(f, 'ERR', e)
P
