#!/bin/bash
timeout 300 python profiles/micro_train_phases.py > gpurun_out/train_phases.txt 2>&1; echo rc=$?; cat gpurun_out/train_phases.txt
timeout 300 python profiles/micro_train_phases.py 32 > gpurun_out/train_phases32.txt 2>&1; echo rc=$?; cat gpurun_out/train_phases32.txt
