"""Markdown table of per-launch headline metrics from an .ncu-rep: python tools/ncu_table.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]
def col(r, k):
    return r[hdr.index(k)] if k in hdr else ""
print("| # | kernel | duration us | DRAM read MB | DRAM write MB | tensor pipe active % | issue active % | L1 % | DRAM % | regs |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
tot_b = tot_t = 0.0
def to_mb(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
units = rows[1]
for i, r in enumerate(rows[2:]):
    name = col(r, "Kernel Name").split("(")[0].split("::")[-1]
    du = float(col(r, "gpu__time_duration.sum").replace(",", ""))
    uu = units[hdr.index("gpu__time_duration.sum")]
    du = du / 1e3 if uu in ("ns", "nsecond") else du * 1e3 if uu in ("ms", "msecond") else du
    rd = to_mb(col(r, "dram__bytes_read.sum"), units[hdr.index("dram__bytes_read.sum")])
    wr = to_mb(col(r, "dram__bytes_write.sum"), units[hdr.index("dram__bytes_write.sum")])
    tot_b += rd + wr; tot_t += du
    print(f"| {i} | `{name}` | {du:.1f} | {rd:.0f} | {wr:.0f} | {col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')[:5]} | "
          f"{col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')[:5]} | {col(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed')[:5]} | "
          f"{col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[:5]} | {col(r, 'launch__registers_per_thread')} |")
n = len(rows) - 2
print(f"\naverage DRAM traffic per launch: {tot_b / n:.1f} MB; average duration {tot_t / n:.1f} us")
