#!/bin/bash
# ncu --set full over the kernels of ONE steady-state forward (final build): GEMM forms, recurrence, memory-bound module kernels
export LANES=8
python profiles/micro_fwd.py > gpurun_out/fwd_plain_r3c.log 2>&1 && ncu --set full --clock-control none -k regex:'gemm_tcgen05|lstm_fused|cos_inst|layernorm|sum_T|ff_attn|attnvideo|rowdot|existsframe|temporal_relate|word_embed|stage_rows' -s 310 -c 62 -o gpurun_out/r2_fwd_full python profiles/micro_fwd.py > gpurun_out/ncu_r3c.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2_fwd_full.ncu-rep --page raw --csv > gpurun_out/r2_fwd_full_raw.csv 2>/dev/null; ls -la gpurun_out/r2_fwd_full* | head; cat gpurun_out/fwd_plain_r3c.log
rm -f gpurun_out/r2_fwd_full.ncu-rep
