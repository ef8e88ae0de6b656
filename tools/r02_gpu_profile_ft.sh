#!/bin/bash
# ncu launch list of one full fine-tune step (BASELINE config 5, ViT-B/16, batch 256): per-kernel time shares.
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
CMD="python tools/ft_profile_step.py 256"
timeout 200 $CMD > gpurun_out/ft_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ft_plain.log; exit 1; }
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 1400 --csv --log-file gpurun_out/ft_launches.csv $CMD > gpurun_out/ncu_ft_launch.log 2>&1
echo "ft launch-list rc=$?"; wc -l gpurun_out/ft_launches.csv
