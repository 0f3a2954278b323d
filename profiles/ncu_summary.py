"""Key counters of every launch in an ncu report, as a table (what the files profiles/r2_ncu_*.txt hold).
  python profiles/ncu_summary.py gpurun_out/<name>.ncu-rep [kernel-regex]"""
import csv, io, re, subprocess, sys

COLS = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "rdGB"), ("dram__bytes_write.sum", "wrGB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1wave%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "longsb"),
        ("launch__grid_size", "grid")]
SCALE = {"us": 1e-3, "s": 1e3, "ns": 1e-6, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
print("%-58s" % "kernel", " ".join("%8s" % c[1] for c in COLS))
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if pat and not pat.search(name):
        continue
    out = []
    for c, _ in COLS:
        if c not in idx:
            out.append("%8s" % "-")
            continue
        try:
            x = float(r[idx[c]].replace(",", "")) * SCALE.get(units[idx[c]], 1.0)
            out.append("%8.3f" % x if x < 1e5 else "%8.3g" % x)
        except ValueError:
            out.append("%8s" % r[idx[c]][:8])
    print("%-58s" % name[:58], " ".join(out))
