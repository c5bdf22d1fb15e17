"""Compact text summary of an `ncu --set full` report (one line block per captured launch).

  python tools/ncu_summary.py gpurun_out/conv2_full.ncu-rep > profiles/rNN_conv2_ncu_full.txt
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (active)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__cycles_active.avg", "smsp cycles active"),
    ("sm__cycles_elapsed.max", "sm cycles elapsed"),
    ("smsp__inst_executed.sum", "instructions"),
]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in out.splitlines() if not l.startswith("==")))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print("# ncu --set full --clock-control none: %s (%d launches)" % (sys.argv[1].split("/")[-1], len(data)))
for r in data:
    print("\n%s  [launch id %s]" % (r[col["Kernel Name"]], r[col["ID"]]))
    for k, label in KEYS:
        if k in col:
            print("  %-34s %14s %s" % (label, r[col[k]], units[col[k]]))
