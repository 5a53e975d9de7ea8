#!/bin/bash
timeout 300 python profiles/module_timeline.py 4096 full > gpurun_out/module_timeline_full.txt 2>&1; echo rc=$?
timeout 300 python profiles/module_timeline.py 4096 modules > gpurun_out/module_timeline_modules.txt 2>&1; echo rc=$?
cat gpurun_out/module_timeline_full.txt
