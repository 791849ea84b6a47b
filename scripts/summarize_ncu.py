"""Turn gpurun_out/launches_<tag>.csv + gpurun_out/prof_<tag>.ncu-rep into tracked summaries under profiles/.

    python scripts/summarize_ncu.py <tag>

Writes profiles/<tag>_launches.csv (kernel, launches, mean us, share of our kernels' time),
profiles/<tag>_ncu_summary.md (key counters of the fwd / bwd kernels) and updates profiles/traffic.json
(dram bytes per launch, read by bench.py for roofline.traffic)."""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"void |fasn::<unnamed>::|at::<unnamed>::|at::native::", "", name)[:70]

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault(short(r[4]), []).append(float(r[-1].replace(",", "")) / 1e3)
ours = sum(sum(v) for k, v in agg.items() if "fasn" in k)
with open(os.path.join(out, f"{tag}_launches.csv"), "w") as f:
    f.write("kernel,launches,mean_us,share_of_fasn_time_pct\n")
    for k, v in agg.items():
        f.write(f"\"{k}\",{len(v)},{sum(v) / len(v):.1f},{100 * sum(v) / ours if 'fasn' in k else 0:.1f}\n")

# ---- full capture
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
want = [("gpu__time_duration.sum", "duration"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "DRAM read % of peak"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "LSU shared wavefronts %"),
        ("launch__registers_per_thread", "registers/thread"), ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue")]
traffic_path = os.path.join(out, "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
with open(os.path.join(out, f"{tag}_ncu_summary.md"), "w") as f:
    f.write(f"# ncu --set full --clock-control none, tag {tag}: bench.py --steps 2 --warmup 3 (C3: fwd+bwd fp16 B4 H32 S4096 D128 n=0.5 causal dropout 0.1)\n\n")
    f.write("Per-launch times under the profiler are cold-cache and serialised: compare shares, not absolutes.\n\n")
    for r in rr[2:]:
        name = short(r[hdr.index("Kernel Name")])
        f.write(f"## {name}\n\n| counter | value | unit |\n|---|---|---|\n")
        vals = {}
        for key, label in want:
            if key in hdr:
                i = hdr.index(key)
                vals[key] = r[i]
                f.write(f"| {label} (`{key}`) | {r[i]} | {units[i]} |\n")
        f.write("\n")
        try:
            rd, wr = float(vals["dram__bytes_read.sum"]), float(vals["dram__bytes_write.sum"])
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
            kind = "bwd_main" if "bwd_kernel" in name else "fwd" if "fwd_kernel" in name else None
            if kind:
                traffic.setdefault("c3", {})[kind + "_dram_bytes"] = (rd + wr) * scale
        except Exception:
            pass
if "c3" in traffic and "bwd_main_dram_bytes" in traffic["c3"]:
    traffic["c3"]["dram_bytes_per_launch"] = traffic["c3"]["bwd_main_dram_bytes"]
    traffic["c3"]["source"] = f"profiles/{tag}_ncu_summary.md"
json.dump(traffic, open(traffic_path, "w"), indent=1)
print(open(os.path.join(out, f"{tag}_launches.csv")).read())
