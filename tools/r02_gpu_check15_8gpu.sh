#!/bin/bash
# Round 2: headline config at N = 1, 2, 4, 8 on ONE box, lean legs, with the duration of the trainable half per N.
mkdir -p gpurun_out
python -m vlm_clip_b200.build > /dev/null 2>&1
timeout 200 python bench.py --gpus 1 --steps 50 --warmup 3 --lean --no-cpu-baseline --no-full-finetune > gpurun_out/r02_scale_1.json 2> gpurun_out/r02_scale_1.err; echo "N=1 rc=$?"
for n in 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 50 --warmup 3 --lean > gpurun_out/r02_scale_$n.json 2> gpurun_out/r02_scale_$n.err; echo "N=$n rc=$?"
done
for n in 1 2 4 8; do python - <<PY
import json
d=json.loads(open("gpurun_out/r02_scale_$n.json").read())
print($n, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "tail", d.get("tail") and {k:round(v,3) for k,v in d["tail"].items() if isinstance(v,float)}, d["clocks"]["sm_mhz"])
PY
done
