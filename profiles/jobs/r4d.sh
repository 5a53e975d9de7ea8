#!/bin/bash
# final evidence of session 3 (length-sorted text recurrence): GPU suite, smoke, default bench, ncu launch list of the bench command
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r3_gpu_tests.log; cat gpurun_out/r3_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r3_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r3_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r3_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-train > gpurun_out/r3_ncu_bench.log 2>&1; echo "ncu rc=$?"
python profiles/launch_summary.py gpurun_out/r3_launches_bench.csv 4 > gpurun_out/r3_launch_summary_bench.txt 2>&1; head -12 gpurun_out/r3_launch_summary_bench.txt
