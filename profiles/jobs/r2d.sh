#!/bin/bash
# first run of the weight-stationary recurrence and of the MN-major CTA-pair GEMM, then the full bench
python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "tn_matches" > gpurun_out/gemm_tn_r2d.log 2>&1; echo "tn tests rc=$?"; tail -4 gpurun_out/gemm_tn_r2d.log
export STAIR_LSTM_WS=1
timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "fused_lstm or full_size" > gpurun_out/ws_tests_r2d.log 2>&1; WS_RC=$?; echo "ws tests rc=$WS_RC"; tail -15 gpurun_out/ws_tests_r2d.log
timeout 300 python profiles/micro_lstm_ws.py > gpurun_out/micro_lstm_ws_r2d.txt 2>&1; echo "micro rc=$?"; cat gpurun_out/micro_lstm_ws_r2d.txt
unset STAIR_LSTM_WS
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r2d.err
python profiles/jobs/summarize_bench.py gpurun_out/bench_r2d.json
if [ "$WS_RC" = "0" ]; then
  export STAIR_LSTM_WS=1
  python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2d_ws.log 2>&1; echo "full pytest (ws) rc=$?"; tail -6 gpurun_out/gpu_tests_r2d_ws.log
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_r2d_ws.json 2> gpurun_out/bench_r2d_ws.err; echo "bench ws rc=$?"
  python profiles/jobs/summarize_bench.py gpurun_out/bench_r2d_ws.json | cut -c1-1200
fi
