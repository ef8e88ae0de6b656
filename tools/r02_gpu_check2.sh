#!/bin/bash
# Round 2, second GPU pass: new trainer tests (graph / resume / stale packs), then the rewritten bench (graph + eager).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -x -k "graph or resume or eval_after or model_m_forward or trainer_step" > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
tail -30 gpurun_out/r02_pytest2.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench2.err; cat gpurun_out/r02_bench2.json
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-full-finetune > gpurun_out/r02_bench2_100.json 2> gpurun_out/r02_bench2_100.err; echo "bench100 rc=$?"
cat gpurun_out/r02_bench2_100.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench2_ref.json 2> gpurun_out/r02_bench2_ref.err; echo "ref rc=$?"
cat gpurun_out/r02_bench2_ref.json
