#!/bin/bash
# streaming recurrence with 16 gate-serial epilogue warps vs the 8-warp product form
timeout 300 python profiles/micro_lstm_ws.py > gpurun_out/micro_lstm_r2i.txt 2>&1; echo "micro rc=$?"; cat gpurun_out/micro_lstm_r2i.txt
