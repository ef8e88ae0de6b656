"""profiles/r01_attention_split.txt from gpurun_out/ (after tools/gpu_split_check.sh and tools/gpu_split_ncu.sh)."""
import csv, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import summarize_profiles as SP

tag = SP.tag
name = f"{tag}_attention_split.txt"
hdr, units, data = SP.report("split_attn.ncu-rep", name,
                             "tcgen05 attention, key-range split (ViT-L/14: B=512, S=257, H=16), launches 6 and 7 of "
                             "`python tools/attn_only.py 512 257 16`", ["keys [0,208)", "keys [208,257)"])
out = (SP.OUT / name).read_text()
lines = [l for l in open(SP.G / "split_launches.csv") if not l.startswith("==")]
per = {}
for d in csv.DictReader(lines):
    k = "first range <0,1>" if "<0, 1>" in d["Kernel Name"] else "second range <0,3>"
    per.setdefault(k, []).append(float(d["Metric Value"]) / 1000)
out += "\n# launch list (ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attention -c 12; serialised)\n"
for k, v in per.items():
    out += f"{k:24s} n={len(v)} avg {sum(v) / len(v):7.1f} us  min {min(v):7.1f}  max {max(v):7.1f}\n"
out += "\n# live timing of the variants (CUDA events, 10 calls after 3 warm-up; VLMCLIP_ATTN_DEBUG phase timers of CTA 0)\n"
out += (SP.G / "split_ab.log").read_text()
(SP.OUT / name).write_text(out)
(SP.OUT / f"{tag}_launches_attention_split.csv").write_text("".join(lines))
print(out)
