#!/bin/bash
# Round 2, first GPU pass: all GPU tests, smoke, residual-GEMM configurations, short bench with both residual modes.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -40 gpurun_out/r02_pytest1.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke1.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke1.log
tail -3 gpurun_out/r02_smoke1.log
for c in 52 43 61; do VLMCLIP_GEMM_RES2_CFG=$c timeout 300 python tools/res2_bench.py; done > gpurun_out/r02_res2_bench.log 2>&1
cat gpurun_out/r02_res2_bench.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-full-finetune > gpurun_out/r02_bench_hilo.json 2> gpurun_out/r02_bench_hilo.err; echo "bench hilo rc=$?"
VLMCLIP_RESIDUAL=bf16 timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-full-finetune > gpurun_out/r02_bench_bf16.json 2> gpurun_out/r02_bench_bf16.err; echo "bench bf16 rc=$?"
tail -2 gpurun_out/r02_bench_hilo.err; cat gpurun_out/r02_bench_hilo.json; cat gpurun_out/r02_bench_bf16.json
