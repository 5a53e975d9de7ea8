#!/bin/bash
timeout 300 python profiles/micro_i3d_phases.py > gpurun_out/i3d_phases.txt 2>&1; echo rc=$?; cat gpurun_out/i3d_phases.txt
