"""Print the interesting keys of a bench.py JSON line (gpurun job helper)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ('value', 'ms_per_step', 'launches_per_step', 'host_enqueue_ms', 'e2e', 'roofline', 'phases_ms', 'strict', 'audit', 'i3d', 'parity', 'cpu_baseline'):
    print(k, json.dumps(d.get(k))[:1600])
print('train', json.dumps(d.get('train'))[:1000])
print('hbm', json.dumps(d.get('roofline_hbm'))[:3200])
