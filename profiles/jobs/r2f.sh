#!/bin/bash
# second look at the weight-stationary recurrence (relaxed spin, L2 prefetch warp, global-space proxy fence) + pair-MN A/B on the train step
export STAIR_LSTM_WS=1
timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "fused_lstm or full_size" > gpurun_out/ws_tests_r2f.log 2>&1; echo "ws tests rc=$?"; tail -3 gpurun_out/ws_tests_r2f.log
timeout 300 python profiles/micro_lstm_ws.py > gpurun_out/micro_lstm_ws_r2f.txt 2>&1; echo "micro rc=$?"; cat gpurun_out/micro_lstm_ws_r2f.txt
unset STAIR_LSTM_WS
for cfg in 'PAIR=0' 'PAIR=1 PAIR_MN=0' 'PAIR=1 PAIR_MN=1'; do echo "== $cfg"; env $cfg python profiles/micro_train.py 6 2>&1 | tail -3; done
