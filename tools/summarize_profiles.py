"""Turns the ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (run here, after tools/gpu_profile.sh)."""
import collections, csv, json, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
G = ROOT / "gpurun_out"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

def launch_summary():
    lines = [l for l in open(G / "launches.csv") if not l.startswith("==")]
    rows = []
    for d in csv.DictReader(lines):
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = d["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        rows.append((re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("vlmclip::<unnamed>::", "")[:72], v))
    agg = collections.OrderedDict()
    for n, v in rows:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    cmd = ("VLMCLIP_OVERLAP_TOWERS=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-finetune --no-graph --lean "
           "--settle-s 0" if tag != "r01" else "VLMCLIP_OVERLAP_TOWERS=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
    skip = 1200 if tag != "r01" else 700
    out = [f"# {tag} ncu launch list: `{cmd}`,",
           f"# ncu --metrics gpu__time_duration.sum --clock-control none -s {skip} -c 520 (about 2.3 steps; cold-cache and",
           f"# serialised: compare SHARES, not absolutes).  raw CSV: {tag}_launches_bench.csv", f"total {tot:.0f} us over {len(rows)} launches", ""]
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{a[1]:10.1f} us {100*a[1]/tot:5.1f}% n={a[0]:4d} avg={a[1]/a[0]:8.1f}  {n}")
    (OUT / f"{tag}_launch_summary.txt").write_text("\n".join(out) + "\n")
    (OUT / f"{tag}_launches_bench.csv").write_text("".join(lines))

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU/F2FP) pipe % of peak"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"), ("launch__cluster_dim_x", "cluster x"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall long scoreboard %"),
]

def raw(rep):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return rows[0], rows[1], rows[2:]

def report(rep, name, title, labels):
    hdr, units, data = raw(G / rep)
    out = [f"# {tag} {title}", f"# ncu --set full --clock-control none --import-source on ({rep}); one column per captured launch",
           "# (ncu's HMMA / tensor-pipe counters do not count tcgen05.mma on this driver, so tensor utilisation is taken from",
           "#  FLOPs / duration in bench.py's roofline, not from ncu)", ""]
    out.append(f"{'metric':48s} " + " ".join(f"{l:>16s}" for l in labels))
    kn = hdr.index("Kernel Name")
    out.append(f"{'kernel':48s} " + " ".join(f"{re.sub(r'.*::', '', d[kn].split('(')[0])[-16:]:>16s}" for d in data))
    for m, nice in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        out.append(f"{nice + ' [' + units[i] + ']':48s} " + " ".join(f"{d[i][:16]:>16s}" for d in data))
    (OUT / name).write_text("\n".join(out) + "\n")
    return hdr, units, data

def mb(v, u): return float(v.replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]


if __name__ == "__main__":
    launch_summary()
    hdr, units, data = report("prof_gemm.ncu-rep", f"{tag}_gemm_layer.txt",
                              "tcgen05 GEMM, the four launches of vision layer 5 (ViT-B/16, B=256, M=50432)",
                              ["qkv(fold)", "out(+res)", "fc1(fold+gelu)", "fc2(+res)"])
    r, w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    per = [mb(d[r], units[r]) + mb(d[w], units[w]) for d in data]
    json.dump({"kernel": "gemm_bf16_tn_kernel", "source": f"profiles/{tag}_gemm_layer.txt (ncu --set full, vision layer 5)",
               "launches": ["qkv", "out_proj", "fc1", "fc2"], "dram_bytes": per,
               "dram_bytes_per_launch": sum(per) / len(per)}, open(OUT / "gemm_traffic.json", "w"), indent=1)
    report("prof_attn.ncu-rep", f"{tag}_attention_pp.txt", "tcgen05 ping-pong attention, vision (B=256, S=197, H=12)", ["vision"])
    if (G / "prof_loss.ncu-rep").exists():
        report("prof_loss.ncu-rep", f"{tag}_clip_loss_strips.txt",
               "strip-wise contrastive loss at N = 4096, P = 768, 512 local rows (tools/loss_only.py)",
               ["norm2", "strip_lse", "strip_grad"])
    print("ok")
