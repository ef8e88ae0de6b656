#!/bin/bash
# ncu capture (SpeedOfLight + memory + launch sections) of the backbone-backward kernels of one full fine-tune step
# (BASELINE config 5, ViT-B/16, batch 128).  Only the raw CSV travels back.
mkdir -p gpurun_out
CMD="python tools/ft_profile_step.py 128"
$CMD > gpurun_out/ft_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ft_plain.log; exit 1; }
# per layer (reverse order): 12 transposes, 4 rowsums, 1 gelu_bwd, 2 LN bwd, 2 attention_bwd (dq + dkv): skip the first
# step's launches (24 layers x 23) and take two layers of the second
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none \
    -k regex:"attention_bwd_dq|attention_bwd_dkv|transpose64|layernorm_bwd_kernel|quick_gelu_bwd|rowsum_bf16" \
    -s 560 -c 48 -o /tmp/prof_ft $CMD > gpurun_out/ncu_ft.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_ft.ncu-rep --page raw --csv > gpurun_out/prof_ft_raw.csv 2>/dev/null
ls -la gpurun_out/prof_ft_raw.csv
