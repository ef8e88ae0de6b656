#!/bin/bash
# Round 2, seventh GPU pass (1 GPU): fused LN statistics, graph edge types (is PDL kept under capture?), A/B benches.
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "finalises or gemm_res2 or test_gemm" > gpurun_out/r02_pytest7a.log 2>&1; echo "pytest7a rc=$?" >> gpurun_out/r02_pytest7a.log
tail -6 gpurun_out/r02_pytest7a.log
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_fullsize.py -m gpu -q --tb=short -x > gpurun_out/r02_pytest7b.log 2>&1; echo "pytest7b rc=$?" >> gpurun_out/r02_pytest7b.log
tail -6 gpurun_out/r02_pytest7b.log
timeout 60 python tools/attn_only.py 256 197 12 > gpurun_out/r02_attn7.log 2>&1; timeout 60 python tools/attn_only.py 512 257 16 >> gpurun_out/r02_attn7.log 2>&1; cat gpurun_out/r02_attn7.log
VLMCLIP_GRAPH_DUMP=gpurun_out/r02_graph timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-full-finetune > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench7.err; cat gpurun_out/r02_bench7.json
for f in gpurun_out/r02_graph.towers.dot gpurun_out/r02_graph.heads.dot; do echo "$f: nodes $(grep -c 'label' $f) programmatic $(grep -ci 'programmatic' $f)"; done
grep -i "programmatic" gpurun_out/r02_graph.towers.dot | head -3
head -c 1500 gpurun_out/r02_graph.towers.dot
rm -f gpurun_out/r02_graph.towers.dot gpurun_out/r02_graph.heads.dot
VLMCLIP_FUSED_STATS=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-full-finetune --lean > gpurun_out/r02_bench7_nofuse.json 2> gpurun_out/r02_bench7_nofuse.err; echo "bench nofuse rc=$?"
cat gpurun_out/r02_bench7_nofuse.json
VLMCLIP_PDL=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-full-finetune --lean --no-graph > gpurun_out/r02_bench7_nopdl_eager.json 2> gpurun_out/r02_bench7_nopdl_eager.err; echo "bench nopdl eager rc=$?"
cat gpurun_out/r02_bench7_nopdl_eager.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-full-finetune --lean --no-graph > gpurun_out/r02_bench7_eager.json 2> gpurun_out/r02_bench7_eager.err; echo "bench eager rc=$?"
cat gpurun_out/r02_bench7_eager.json
