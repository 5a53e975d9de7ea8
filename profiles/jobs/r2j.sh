#!/bin/bash
# PROBE: what each class of epilogue memory/MUFU work costs in the streaming recurrence (results are wrong by construction)
for sk in 0 1 2 4 8 3 7 15; do
  echo "skip=$sk"; STAIR_LSTM_SKIP=$sk timeout 200 python profiles/micro_lstm_one.py 2>&1 | tail -2
done
