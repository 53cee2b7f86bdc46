#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/xyz.md"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (of elapsed)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n")
    print(f"Source: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`), one column per "
          "captured launch.\n")
    names = [r[hdr.index("Kernel Name")].split("(")[0][-40:] for r in rows[2:]]
    print("| metric | " + " | ".join(f"#{i} {n}" for i, n in enumerate(names)) + " |")
    print("|---|" + "---:|" * len(names))
    for key, label in METRICS:
        if key not in hdr:
            continue
        j = hdr.index(key)
        vals = []
        for r in rows[2:]:
            try:
                vals.append(f"{float(r[j].replace(',', '')):.2f}")
            except ValueError:
                vals.append(r[j])
        print(f"| {label} [{units[j]}] | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
