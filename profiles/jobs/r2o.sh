#!/bin/bash
timeout 120 python profiles/micro_gemm_timeline.py > gpurun_out/gemm_timeline.txt 2>&1; echo rc=$?; cat gpurun_out/gemm_timeline.txt
