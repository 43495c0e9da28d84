"""Turn gpurun_out/launches_<tag>.csv and prof_<tag>.ncu-rep into the small CSVs kept under profiles/."""
import csv
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "final"
rnd = sys.argv[2] if len(sys.argv) > 2 else "r1"

rows = [r for r in csv.reader(l for l in open("gpurun_out/launches_%s.csv" % tag) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg, tot = {}, 0.0
for r in rows[1:]:
    n = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    tot += v
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
with open("profiles/%s_launches_%s.csv" % (rnd, tag), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_us", "share_pct"])
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        w.writerow([n, c, round(v / 1000, 1), round(100 * v / tot, 2)])
        print("%-60s %4d %10.1f us %5.1f%%" % (n[:60], c, v / 1000, 100 * v / tot))

raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_%s.ncu-rep" % tag, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
keep = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
idx = [hdr.index(k) for k in keep if k in hdr]
with open("profiles/%s_ncu_sections_%s.csv" % (rnd, tag), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([rows[1][i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
        print(r[hdr.index("Kernel Name")][:48], {hdr[i].split(".")[0][-28:]: r[i][:12] for i in idx[3:10]})
