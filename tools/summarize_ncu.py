"""Summarise an .ncu-rep (ncu --set full) into a compact per-launch CSV + markdown table for profiles/."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "dur_us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("sm__cycles_elapsed.max.per_second", "sm_ghz"), ("launch__occupancy_limit_shared_mem", "occ_limit_smem")]
units = rows[1]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([n for _, n in want])
    for r in data:
        line = []
        for h, n in want:
            v = r[col[h]] if h in col else ""
            if n == "kernel":
                v = v.split("(")[0].replace("void nb200::<unnamed>::", "")
            if n == "dur_us" and v:
                u = units[col[h]]   # ncu picks the unit per report: ns / us / ms / s
                v = "%.3f" % (float(v) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(u, 1.0))
            if n in ("dram_rd_MB", "dram_wr_MB") and v:
                u = units[col[h]]
                v = "%.2f" % (float(v) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0))
            line.append(v)
        w.writerow(line)
print(open(out).read())
