"""Condense `ncu -i k1.ncu-rep --page raw --csv` (tools/ncu_target_k1.py, crop_kernel + pack_u8_kernel, --set full) into
profiles/rNN_ncu_k1_b32.csv.   python tools/ncu_k1_summarise.py raw.csv out.csv"""
import csv, io, sys
rows = list(csv.DictReader(io.StringIO("".join(l for l in open(sys.argv[1]) if l.startswith('"')))))
units, rows = rows[0], rows[1:]
COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct"]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(COLS)
    w.writerow([units.get(c, "") for c in COLS])
    for r in rows:
        w.writerow([r.get(c, "") for c in COLS])
print("%d launches" % len(rows))
