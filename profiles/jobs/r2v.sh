#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -4
for pg in 0 1; do
python - <<PY
import sys, os
sys.path.insert(0, '.')
from stair_b200 import _lib as L
L.lib().stair_set_gemm_pair_gather($pg)
print('pair gather $pg')
sys.argv = ['x']
exec(open('profiles/micro_i3d_phases.py').read().split("lib.stair_debug_timeline(1)")[0])
PY
done 2>&1 | grep -v "^$" > gpurun_out/pair_gather_ab.txt; cat gpurun_out/pair_gather_ab.txt
for pg in 0 1; do
python - <<PY
import sys, os
sys.path.insert(0, '.')
from stair_b200 import _lib as L
L.lib().stair_set_gemm_pair_gather($pg)
print('pair gather $pg (RX)')
sys.argv = ['x']
exec(open('profiles/micro_fwd.py').read())
PY
done 2>&1 | grep -v "^$" >> gpurun_out/pair_gather_ab.txt; tail -4 gpurun_out/pair_gather_ab.txt
