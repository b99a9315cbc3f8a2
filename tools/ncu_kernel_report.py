"""Per-kernel roofline table from one or more .ncu-rep files (ncu --set full):

  python tools/ncu_kernel_report.py gpurun_out/r02_a.ncu-rep [more.ncu-rep ...] > profiles/r02_kernels.md

For every kernel name: launches captured, mean duration, DRAM bytes read+written per launch, achieved DRAM GB/s and
its fraction of the measured copy peak (MEASURED_PEAKS.json), tensor-pipe / XU (MUFU) / FMA / ALU pipe utilisation,
issue-slot utilisation, shared-memory throughput, achieved occupancy, registers, and the top warp-stall reasons."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK_GB = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK_GB = 6650.0

M = {
    "dur": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor2": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "alu": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smem": "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "l1": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "occ": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
}
STALLS = {k: f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio" for k in (
    "long_scoreboard", "short_scoreboard", "mio_throttle", "lg_throttle", "barrier", "math_pipe_throttle", "wait",
    "tex_throttle", "membar", "sleeping", "dispatch_stall", "no_instruction", "not_selected", "selected", "imc_miss",
    "drain", "branch_resolving")}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
         "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}


def load(path):
    """.ncu-rep (converted through `ncu -i ... --page raw --csv`) or an already converted raw-page .csv"""
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def val(hdr, units, r, key):
    if key not in hdr:
        return None
    i = hdr.index(key)
    try:
        return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
    except ValueError:
        return None


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    name = re.sub(r"lc::\(anonymous namespace\)::|lc::", "", name)
    return name


def main(paths):
    agg = collections.OrderedDict()
    for p in paths:
        hdr, units, rows = load(p)
        for r in rows:
            name = short(r[hdr.index("Kernel Name")])
            grid = r[hdr.index(M["grid"])] if M["grid"] in hdr else ""
            d = agg.setdefault((name, grid), collections.defaultdict(list))
            for k, m in M.items():
                v = val(hdr, units, r, m)
                if v is not None:
                    d[k].append(v)
            for k, m in STALLS.items():
                v = val(hdr, units, r, m)
                if v is not None:
                    d["stall_" + k].append(v)
    mean = lambda xs: sum(xs) / len(xs) if xs else float("nan")  # noqa: E731
    print(f"HBM denominator: {PEAK_GB:.1f} GB/s (MEASURED_PEAKS.json copy bandwidth).  `ncu --set full --clock-control none`; "
          "durations are ncu's serialised single-launch times (cold L2), so GB/s here is a lower bound of the in-step rate.\n")
    print("| kernel (grid) | n | us | DRAM MB rd+wr | GB/s | of HBM peak | DRAM % | tensor % | XU % | FMA % | ALU % | issue % | smem % | L1 % | L2 % | occ % | regs |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for (name, grid), d in agg.items():
        us = mean(d["dur"])
        mb = (mean(d["rd"]) + mean(d["wr"])) / 1e6
        gbs = mb * 1e6 / (us * 1e-6) / 1e9 if us > 0 else 0
        f = lambda k: f"{mean(d[k]):.1f}" if d[k] else "-"  # noqa: E731
        tensor = f("tensor") if d["tensor"] else f("tensor2")
        print(f"| `{name}` ({grid}) | {len(d['dur'])} | {us:.1f} | {mb:.1f} | {gbs:.0f} | {gbs / PEAK_GB:.2f} | {f('dram_pct')} | "
              f"{tensor} | {f('xu')} | {f('fma')} | {f('alu')} | {f('issue')} | {f('smem')} | {f('l1')} | {f('l2')} | {f('occ')} | "
              f"{int(mean(d['regs'])) if d['regs'] else '-'} |")
    print("\nTop warp-stall reasons (warps stalled per issue-active cycle):\n")
    for (name, grid), d in agg.items():
        st = sorted(((mean(v), k[6:]) for k, v in d.items() if k.startswith("stall_") and v), reverse=True)[:4]
        print(f"- `{name}` ({grid}): " + ", ".join(f"{k} {v:.2f}" for v, k in st))


if __name__ == "__main__":
    main(sys.argv[1:])
