#!/bin/bash
# ncu captures (SpeedOfLight + memory sections) of the HBM / latency-bound kernels of one step: row kernels, adapters,
# loss, optimizer.  Only the raw CSV travels back (the .ncu-rep of ~50 launches is larger than gpurun's copy-back limit).
mkdir -p gpurun_out
export VLMCLIP_OVERLAP_TOWERS=0
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > /dev/null 2>&1 || { echo "plain run failed"; exit 1; }
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none \
    -k regex:"layernorm_kernel|vision_embed_ln|ln_partials|im2col|text_embed|sim_logits|row_lse|col_lse|clip_grad|clip_norm|clip_loss|l2norm|adapter_fwd|adapter_bwd|adapter_wgrad|adapter_colsum|linear_f32|adamw|sumsq|attention_fwd_kernel" \
    -s 140 -c 48 -o /tmp/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
echo "small rc=$?"
ncu -i /tmp/prof_small.ncu-rep --page raw --csv > gpurun_out/prof_small_raw.csv 2>/dev/null
ls -la gpurun_out/prof_small_raw.csv
