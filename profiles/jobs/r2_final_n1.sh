#!/bin/bash
# final build: full GPU suite, smoke, full bench (N = 1), reference arm
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_final.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/gpu_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final_n1.err
python profiles/jobs/summarize_bench.py gpurun_out/bench_final_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null; echo "ref rc=$?"; cat gpurun_out/bench_final_ref.json | cut -c1-600
