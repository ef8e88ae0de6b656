#!/bin/bash
# Round 2 ncu evidence (each capture only after the same command exited 0 without ncu) + the N = 1 lines of configs 3 / 5.
# Towers serialised and eager launches so that launch indices are deterministic: per step 48 text GEMMs, the patch GEMM,
# then 4 GEMMs per vision layer.
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
export VLMCLIP_OVERLAP_TOWERS=0
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-finetune --no-graph --lean --settle-s 0"
timeout 200 $CMD > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err || { echo "plain run failed"; tail -5 gpurun_out/plain_bench.err; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 520 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch-list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 263 -c 4 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_pp -s 40 -c 1 -o gpurun_out/prof_attn -f $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn rc=$?"
timeout 100 python tools/loss_only.py > gpurun_out/loss_only.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:clip_ -s 9 -c 3 -o gpurun_out/prof_loss -f python tools/loss_only.py > gpurun_out/ncu_loss.log 2>&1
echo "loss rc=$?"; cat gpurun_out/loss_only.log
ls -la gpurun_out/*.ncu-rep
unset VLMCLIP_OVERLAP_TOWERS
timeout 300 python bench.py --steps 20 --warmup 3 --workload cfg3 --lean > gpurun_out/r02_bench_cfg3_1gpu.json 2> gpurun_out/r02_bench_cfg3_1gpu.err; echo "cfg3 rc=$?"; cat gpurun_out/r02_bench_cfg3_1gpu.json
timeout 300 python bench.py --steps 20 --warmup 3 --workload cfg5 > gpurun_out/r02_bench_cfg5_1gpu.json 2> gpurun_out/r02_bench_cfg5_1gpu.err; echo "cfg5 rc=$?"; cat gpurun_out/r02_bench_cfg5_1gpu.json
timeout 300 python bench.py --steps 50 --warmup 3 --lean > gpurun_out/r02_bench_cfg2_1gpu.json 2> gpurun_out/r02_bench_cfg2_1gpu.err; echo "cfg2 rc=$?"; cat gpurun_out/r02_bench_cfg2_1gpu.json
