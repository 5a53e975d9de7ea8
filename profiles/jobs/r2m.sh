#!/bin/bash
timeout 60 profiles/micro/pdl_chain > gpurun_out/pdl_chain.txt 2>&1; echo rc=$?; cat gpurun_out/pdl_chain.txt
