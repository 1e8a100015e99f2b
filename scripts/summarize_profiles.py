"""Turns gpurun_out/ ncu artefacts into the tracked summaries under profiles/."""
import csv, json, os, re, subprocess, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles"); os.makedirs(out, exist_ok=True)
go = os.path.join(root, "gpurun_out")

# 1. launch list -> per-kernel totals and shares
rows = [r for r in csv.reader(open(os.path.join(go, "launches.csv"))) if len(r) > 14 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    d = per.setdefault(name, [0, 0.0])
    d[0] += 1; d[1] += float(r[14]) / 1e6
total = sum(v[1] for v in per.values())
with open(os.path.join(out, f"{tag}_launches_summary.md"), "w") as f:
    f.write(f"# ncu launch list, `bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e` (cfg3, 1 GPU)\n\n"
            f"`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and\n"
            f"serialised; compare SHARES. {len(rows)} launches, {total:.1f} ms total kernel time "
            f"(generator + 1 warm-up + 2 timed steps).\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (n, ms) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / total:.2f}% |\n")
import shutil
shutil.copy(os.path.join(go, "launches.csv"), os.path.join(out, f"{tag}_launches.csv"))

# 2. full capture of the top kernel -> key metrics
rep = os.path.join(go, "knn_tiles_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, vals = rr[0], rr[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled"]
with open(os.path.join(out, f"{tag}_knn_tiles_tc_ncu.txt"), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:k_knn_tiles_tc -s 1 -c 1 "
            "python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e   (cfg3: 1,306,127 cells, d=50, k=20)\n\n")
    for h, u, v in zip(hdr, rr[1], vals):
        if any(w in h for w in want):
            f.write(f"{h} [{u}] = {v}\n")
# 3. DRAM traffic per launch of the dominant kernel -> profiles/traffic.json (read by bench.py)
def num(name):
    return float(vals[hdr.index(name)].replace(",", ""))
unit_r, unit_w = rr[1][hdr.index("dram__bytes_read.sum")], rr[1][hdr.index("dram__bytes_write.sum")]
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
traffic = num("dram__bytes_read.sum") * mult[unit_r] + num("dram__bytes_write.sum") * mult[unit_w]
json.dump({"_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture per kernel "
                      f"(profiles/{tag}_knn_tiles_tc_ncu.txt), per launch",
           "k_knn_tiles_tc": {"workload": "cfg3", "n_gpus": 1, "bytes_per_launch": int(traffic),
                              "capture": f"profiles/{tag}_knn_tiles_tc_ncu.txt"}},
          open(os.path.join(out, "traffic.json"), "w"), indent=1)
print(open(os.path.join(out, f"{tag}_launches_summary.md")).read()[:3000])
print(open(os.path.join(out, f"{tag}_knn_tiles_tc_ncu.txt")).read()[:5000])
