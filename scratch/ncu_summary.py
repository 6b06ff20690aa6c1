"""profiles/<tag>_ncu_summary.md + profiles/<tag>_launches.csv + profiles/ncu_traffic.json from a capture made by
profiles/run_ncu.sh (gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep).  usage: python scratch/ncu_summary.py <tag> [bench json]"""
import csv, collections, io, json, shutil, subprocess, sys
tag = sys.argv[1]
bench = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else None
rows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    name = r[kn].split("(")[0].replace("void ", "").replace("dns::", "")
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
own = sum(v[1] for k, v in agg.items() if k.startswith("k_"))
out = [f"# ncu evidence, capture `{tag}`", "",
       "Command (B200, one GPU): `bash profiles/run_ncu.sh " + tag + "` -- `python bench.py --steps 2 --warmup 1 --rays-per-gpu 32768 "
       "--no-extra --no-cpu` plain (exit 0), then the launch list (`--metrics gpu__time_duration.sum --clock-control none`, "
       f"`profiles/{tag}_launches.csv`) and `--set full --import-source on` of the large kernels.  Per-launch times under ncu are "
       "cold-cache and serialised: read the SHARES.", "",
       "## Launch list: share of the summed device time", "", "| kernel | launches | us total | share |", "|---|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
    out.append(f"| `{k[:70]}` | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f} % |")
out += [f"| **total** ({len(agg)} kernels; this library's `k_*` kernels: {100 * own / tot:.1f} %) | | **{tot / 1e3:.1f}** | |", ""]
if bench:
    ph = bench["roofline"]["phase_ms_per_step"]; ms = bench["ms_per_step"]
    out += [f"CUDA-event phase timers of the plain `bench.py` run at {bench['config']['rays_per_gpu']} rays per step "
            f"({ms:.2f} ms per step): " + ", ".join(f"{k} {v:.2f} ms ({100 * v / ms:.0f} %)" for k, v in sorted(ph.items(), key=lambda x: -x[1]) if v > 0.05) + ".", ""]
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hh = rr[0]
def col(r, name):
    try: return float(r[hh.index(name)].replace(",", ""))
    except (ValueError, IndexError): return float("nan")
out += ["## `--set full` (one launch each; 32 768 rays = 1 540 096 points)", "",
        "| kernel | time ms | DRAM rd + wr GB | regs | warps active | issue active | L1 thr. | L2 thr. (busiest slice) | L1 hit | L2 hit | top stalls |", "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
traffic = {}
for r in rr[2:]:
    name = r[hh.index("Kernel Name")].split("(")[0].replace("void ", "")
    stalls = []
    for i, c in enumerate(hh):
        if c.startswith("smsp__pcsamp_warps_issue_stalled_") and not c.endswith("_not_issued"):
            try: stalls.append((float(r[i]), c.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError: pass
    ts = sum(v for v, _ in stalls) or 1.0
    top = ", ".join(f"{c} {100 * v / ts:.0f} %" for v, c in sorted(stalls, reverse=True)[:3])
    rd, wr = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
    out.append(f"| `{name}` | {col(r, 'gpu__time_duration.sum'):.3f} | {rd:.2f} + {wr:.2f} | {int(col(r, 'launch__registers_per_thread'))} | "
               f"{col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % | {col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} % | "
               f"{col(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} % | {col(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} % "
               f"({col(r, 'lts__throughput.max.pct_of_peak_sustained_elapsed'):.0f} %) | {col(r, 'l1tex__t_sector_hit_rate.pct'):.1f} % | "
               f"{col(r, 'lts__t_sector_hit_rate.pct'):.1f} % | {top} |")
    traffic.setdefault(name, (rd + wr) * 1e9)
pts = 32768 * 47
tj = {"capture": tag, "points_per_launch": pts, "bytes_per_point": {k: v / pts for k, v in traffic.items() if ("point" in k or "ray" in k) and "<2>" not in k},   # <2>: the TV lattice, not the ray batch
     
      "note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full launch, divided by the launch's sample points"}
json.dump(tj, open("profiles/ncu_traffic.json", "w"), indent=1)
out += ["", "DRAM bytes per sample point (`profiles/ncu_traffic.json`, what `bench.py` scales into `roofline.traffic`): " +
        ", ".join(f"`{k}` {v:.0f} B" for k, v in tj["bytes_per_point"].items()) + "."]
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
shutil.copy(f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launches.csv")
print("\n".join(out))
