#!/bin/bash
# ncu launch list of the bench command itself (quick flags), final build: the first timed forwards of bench.py
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-train"
$CMD > gpurun_out/bench_quick_plain.json 2> gpurun_out/bench_quick_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_bench_final.csv $CMD > gpurun_out/ncu_bench_final.log 2>&1; echo "ncu rc=$?"
python profiles/jobs/summarize_bench.py gpurun_out/bench_quick_plain.json | grep "^value\|^ms_per" 
