"""Condense `ncu -i step.ncu-rep --page raw --csv` of one trunk pass (tools/ncu_target.py, --set full --clock-control none)
into the two tracked artefacts: a per-launch table and the DRAM traffic of the tcgen05 conv launches that bench.py's
`roofline.traffic` quotes.   python tools/ncu_summarise.py raw.csv profiles/r01_ncu_full_step_b32.csv profiles/r01_traffic.json 32"""
import csv, io, json, sys

raw, out_csv, out_json, batch = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
lines = [l for l in open(raw) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
units, rows = rows[0], rows[1:]
COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg"]


def num(r, k):
    return float(str(r[k]).replace(",", ""))


def to_bytes(r, k):
    u = units[k].lower()
    return num(r, k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


def to_ms(r, k):
    u = units[k].lower()
    return num(r, k) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]


with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(COLS)
    w.writerow([units.get(c, "") for c in COLS])
    for r in rows:
        w.writerow([r[c] for c in COLS])
conv = [r for r in rows if any(n in r["Kernel Name"] for n in ("conv_umma_kernel", "conv_rows_kernel", "conv_tsweep_kernel", "stem_sweep_kernel", "conv_bc_fused_kernel"))]
rd = sum(to_bytes(r, "dram__bytes_read.sum") for r in conv)
wr = sum(to_bytes(r, "dram__bytes_write.sum") for r in conv)
allb = sum(to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum") for r in rows)
json.dump({"source": "ncu --set full --clock-control none, tools/ncu_target.py %d (second pass, %d launches), %s" % (batch, len(rows), out_csv),
           "batch": batch, "conv_launches": len(conv), "conv_dram_read_bytes_per_step": rd, "conv_dram_write_bytes_per_step": wr,
           "conv_dram_bytes_per_step": rd + wr,
           "conv_kernel_ms_per_step_under_ncu": sum(to_ms(r, "gpu__time_duration.sum") for r in conv),
           "all_kernels_dram_bytes_per_step": allb}, open(out_json, "w"), indent=1)
print("launches %d, conv launches %d, conv DRAM %.2f GB (read %.2f, write %.2f)" % (len(rows), len(conv), (rd + wr) / 1e9, rd / 1e9, wr / 1e9))
