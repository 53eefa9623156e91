"""Selected raw metrics of every kernel in an .ncu-rep (ncu --set full) as a small CSV for profiles/:
   python scripts/ncu_summary.py <report.ncu-rep> > profiles/<name>_summary.csv"""
import csv, io, subprocess, sys
WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
cols = [c for c in WANT if c in hdr]
w = csv.writer(sys.stdout)
w.writerow(["ID", "Kernel Name"] + cols)
w.writerow(["", ""] + [units[hdr.index(c)] for c in cols])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    w.writerow([d["ID"], d["Kernel Name"]] + [d[c] for c in cols])
