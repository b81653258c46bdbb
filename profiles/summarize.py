"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python profiles/summarize.py launches gpurun_out/launches_r01.csv > profiles/launches_r01_summary.txt
  python profiles/summarize.py full gpurun_out/prof_conv_r01.ncu-rep > profiles/conv_umma_r01_ncu.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0]
        a = agg[name]
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    unit = rows[1][hdr.index("Metric Unit")] if len(rows) > 1 else "?"
    tot = sum(v[1] for v in agg.values())
    print(f"# per-kernel totals over {sum(v[0] for v in agg.values())} profiled launches (ncu gpu__time_duration.sum, unit {unit}; cold-cache, serialised:"
          f" compare SHARES)\n# kernel\tlaunches\ttotal\tshare")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k}\t{n}\t{t:.0f}\t{100 * t / tot:.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "launch__waves_per_multiprocessor",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "dram__cycles_active"]
    idx = [i for i, h in enumerate(hdr) if any(k in h for k in keys) and "per_second" not in h]
    for r in rows[2:]:
        print("----")
        for i in idx:
            print(f"{hdr[i]}\t{r[i]}\t{units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
