#!/bin/bash
# tests + full bench + ncu capture of the CTA-pair GEMM (one ncu family per call)
export STAIR_GRAD_REPORT=gpurun_out/grad_report_r2c.txt; rm -f $STAIR_GRAD_REPORT
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/gpu_tests_r2c.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/gpu_tests_r2c.log
/usr/bin/time -v python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/bench_r2c.err; tail -5 gpurun_out/bench_r2c.err
python profiles/jobs/summarize_bench.py gpurun_out/bench_r2c.json
python profiles/micro_xproj.py > gpurun_out/xproj_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_pair -s 3 -c 1 -o gpurun_out/gemm_pair_r2 python profiles/micro_xproj.py > gpurun_out/ncu_pair.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/xproj_plain.log
python profiles/micro_fwd.py > gpurun_out/fwd_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_fwd_r2c.csv python profiles/micro_fwd.py > gpurun_out/ncu_fwd.log 2>&1; echo "ncu2 rc=$?"; cat gpurun_out/fwd_plain.log
