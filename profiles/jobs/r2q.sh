#!/bin/bash
timeout 300 python profiles/micro_train_host.py 32 > gpurun_out/train_host32.txt 2>&1; echo rc=$?; cat gpurun_out/train_host32.txt
