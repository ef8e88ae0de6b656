#!/bin/bash
# Round 2 final pass (what the driver runs at round end): GPU tests, smoke, bench (native + reference arm).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_final_build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_pytest.log
tail -5 gpurun_out/r02_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_final_smoke.log
tail -2 gpurun_out/r02_final_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err; echo "ref rc=$?"
cut -c1-200 gpurun_out/r02_final_bench_ref.json
/usr/bin/time -v timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
grep "Elapsed (wall" gpurun_out/r02_final_bench.err; cat gpurun_out/r02_final_bench.json
