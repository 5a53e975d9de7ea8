#!/bin/bash
# weight-stationary recurrence form 4 (xproj staged through shared memory by loader warps)
export STAIR_LSTM_WS=1 STAIR_DEBUG=1
timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "fused_lstm or full_size" > gpurun_out/ws_tests_r2h.log 2>&1; echo "ws tests rc=$?"; tail -3 gpurun_out/ws_tests_r2h.log
timeout 300 python profiles/micro_lstm_ws.py > gpurun_out/micro_lstm_ws_r2h.txt 2>&1; echo "micro rc=$?"; cat gpurun_out/micro_lstm_ws_r2h.txt
