#!/usr/bin/env python3
"""Summarise an ncu report (--set full) and a launch list into profiles/ (tracked).

    python tools/summarize_ncu.py gpurun_out/prof_r1_persistent.ncu-rep gpurun_out/launches_r1.csv r1
"""
import csv, json, subprocess, sys, os, collections

rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keep = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "gpc__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
kernels = []
for r in rows[2:]:
    d = {"kernel": r[idx["Kernel Name"]].split("(")[0]}
    for k in keep:
        if k in idx:
            try:
                d[k] = float(r[idx[k]].replace(",", ""))
            except ValueError:
                d[k] = r[idx[k]]
            d[k + " [unit]"] = units[idx[k]]
    kernels.append(d)
with open(os.path.join(out_dir, f"ncu_full_{tag}.json"), "w") as f:
    json.dump(kernels, f, indent=1)
# launch list: per-kernel totals and shares
tot = collections.OrderedDict()
n = collections.Counter()
with open(launches) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
h = next(rd)
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
for r in rd:
    name = r[ki].split("(")[0]
    tot[name] = tot.get(name, 0.0) + float(r[vi].replace(",", ""))
    n[name] += 1
total = sum(tot.values())
with open(os.path.join(out_dir, f"launches_{tag}.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over bench.py\n\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, v in tot.items():
        f.write(f"| {k} | {n[k]} | {v/1e6:.3f} | {100*v/total:.1f}% |\n")
print(open(os.path.join(out_dir, f"launches_{tag}.md")).read())
# per-iteration DRAM traffic of the dominant kernel for bench.py's roofline.traffic
pk = [k for k in kernels if "persistent" in k["kernel"]]
if pk:
    def tobytes(d, key):
        v, u = d[key], d[key + " [unit]"].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    per_launch = sum(tobytes(k, "dram__bytes_read.sum") + tobytes(k, "dram__bytes_write.sum") for k in pk) / len(pk)
    iters = 20  # tools/profile_target.py profiles launches of 20 iterations
    with open(os.path.join(out_dir, "roofline_traffic.json"), "w") as f:
        json.dump({"bgp_persistent_kernel_dram_bytes_per_iteration": per_launch / iters,
                   "source": f"ncu --set full, {os.path.basename(rep)}, launches of {iters} iterations x 256 chains",
                   "note": "the 237 B/eval trace rows are written to L2 and evicted later; they do not show up inside the launch"}, f, indent=1)
    print("dram bytes per iteration:", per_launch / iters)
