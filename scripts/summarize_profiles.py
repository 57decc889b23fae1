"""Turn gpurun_out/ ncu artefacts into the small tracked summaries under profiles/ (run in the build container)."""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
lines = [f"# profiles/{tag} -- ncu summaries (B200, sm_100a)\n"]

launches = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(launches):
    rows = list(csv.reader(open(launches)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hi], rows[hi + 1:]
    ik, iv, ig = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) > iv:
            name = r[ik].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")
            agg[(name[:70], r[ig])].append(float(r[iv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    lines.append(f"## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, python bench.py --steps 2 --warmup 3)\n")
    lines.append("| kernel | grid | launches | mean us | share of GPU time |\n|---|---|---|---|---|")
    for (n, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{n}` | {g} | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {100*sum(v)/tot:.1f} % |")
    lines.append("")

rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, v = r[0], r[2] if len(r) > 2 else r[1]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "sm__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
    lines.append(f"## sqp_kernel, one launch of the bench batch (`ncu --set full --clock-control none`, {os.path.basename(rep)})\n")
    lines.append("| metric | value |\n|---|---|")
    d = dict(zip(h, v))
    u = dict(zip(h, r[1]))
    for k in want:
        if k in d:
            lines.append(f"| {k} | {d[k]} {u.get(k, '')} |")
    lines.append("")
for f in sorted(os.listdir(os.path.join(ROOT, "gpurun_out"))):
    if f.startswith(f"bench_{tag}") and f.endswith(".json"):
        try:
            d = json.loads(open(os.path.join(ROOT, "gpurun_out", f)).read().strip().splitlines()[-1])
            lines.append(f"## {f}\n\n```json\n{json.dumps(d, indent=1)}\n```\n")
        except Exception:
            pass
ph = os.path.join(ROOT, "gpurun_out", f"phases_{tag}.txt")
if os.path.exists(ph):
    lines.append("## Per-phase cycles of one solve (thread 0 clock64, `scripts/profile_phases.py 8`)\n\n```\n" + open(ph).read() + "```\n")
open(os.path.join(out_dir, f"{tag}_summary.md"), "w").write("\n".join(lines))
print("wrote", os.path.join(out_dir, f"{tag}_summary.md"))
