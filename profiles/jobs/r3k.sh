#!/bin/bash
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
python profiles/jobs/summarize_bench.py gpurun_out/bench_default.json | grep "^value\|^ms_per_step\|^train\|^i3d\|^parity" | cut -c1-300
