"""Key metrics of every kernel in an .ncu-rep (one `ncu --set full` capture per kernel) -> a tracked text summary.

    python scripts/summarize_ncu.py gpurun_out/r02_stages_full.ncu-rep profiles/r02_stages_ncu.txt ["header line"]
"""
import csv, subprocess, sys, re

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma", "sm__inst_executed_pipe_tensor_subpipe_imma", "sm__inst_executed_pipe_fp64",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__cycles_elapsed.avg ", "sm__cycles_active.avg ",
]
EXCL = ("TriageCompute", ".max", ".min", "pct_of_peak_sustained_elapsed.", "per_second", ".sum.pct", ".sum.per", "peak_sustained_active.")

def main():
    rep, out = sys.argv[1], sys.argv[2]
    head = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    seen = {}
    with open(out, "w") as f:
        if head:
            f.write(head + "\n\n")
        for row in rr[2:]:
            name = re.sub(r"\(.*", "", row[hdr.index("Kernel Name")]).replace("void ", "")
            seen[name] = seen.get(name, 0) + 1
            if seen[name] > 2:
                continue
            f.write(f"=== {name}  (capture {seen[name]})\n")
            tr = 0.0
            for h, u, v in zip(hdr, units, row):
                hh = h + " "
                if any(k in hh for k in KEYS) and not any(e in h for e in EXCL) and v not in ("", "n/a"):
                    f.write(f"  {h} [{u}] = {v}\n")
                if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tr += float(v.replace(",", "")) * mult.get(u, 1)
            f.write(f"  DRAM traffic per launch (read + write) = {tr / 1e9:.3f} GB\n")
            # warp-state breakdown: the five largest stall reasons per issued instruction
            st = [(float(v.replace(",", "")), h) for h, v in zip(hdr, row)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and v not in ("", "n/a")]
            for val, h in sorted(st, reverse=True)[:5]:
                f.write(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} = {val:.2f} warps/issue\n")
            f.write("\n")
    print(open(out).read())

if __name__ == "__main__":
    main()
