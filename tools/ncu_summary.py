"""Turns an .ncu-rep (ncu --set full) into the per-launch CSV summary kept under profiles/: duration, DRAM bytes, L2 bytes and hit rate,
L1 sectors per request of global loads, issue-active, registers.  Usage: python tools/ncu_summary.py <rep> <out.csv> ["# header line" ...]"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
hdr = sys.argv[3:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
names, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio", "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
idx = [(w, names.index(w)) for w in want if w in names]
with open(out, "w") as f:
    for h in hdr:
        f.write(h.rstrip() + "\n")
    w = csv.writer(f)
    w.writerow(["launch"] + [f"{n} [{units[i]}]" if units[i] else n for n, i in idx])
    for k, r in enumerate(data):
        w.writerow([k] + [r[i][:90] for n, i in idx])
print(open(out).read()[:3000])
