"""Assembles profiles/r02_kernels.md from one bench.py line and the ncu raw pages of tools/gpu_run4.sh / gpu_run9.sh:

  python tools/make_profile_md.py <bench.json> <den.csv> <dec.csv> <gemm.csv> <conv.csv> [gemm_light.csv] > profiles/r02_kernels.md

(.csv = `ncu -i x.ncu-rep --page raw --csv`, gz accepted).  Section 1 is the in-step class table of the bench line
(algorithmic bytes / FLOPs, achieved vs measured peak), sections 2-4 are ncu: per-kernel table of the non-GEMM kernels,
per-kernel + per-launch tables of the tcgen05 GEMM and the implicit-GEMM sphere conv."""
import csv
import gzip
import io
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(path):
    return gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()


def plain(path):
    """ncu_kernel_report.py wants an uncompressed .csv"""
    if not path.endswith(".gz"):
        return path
    f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
    f.write(read(path))
    f.close()
    return f.name


def launch_table(path, title):
    rows = [r for r in csv.reader(io.StringIO("\n".join(l for l in read(path).splitlines() if not l.startswith("=="))))]
    hdr, units, body = rows[0], rows[1], rows[2:]

    def col(r, k):
        return r[hdr.index(k)] if k in hdr else ""

    def to_mb(v, u):
        return float(v.replace(",", "")) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)

    out = [f"### {title}\n", "| # | kernel (grid) | duration us | DRAM read MB | DRAM write MB | tensor pipe active % | issue active % | L2 % | DRAM % |",
           "|---|---|---:|---:|---:|---:|---:|---:|---:|"]
    tb = tt = 0.0
    for i, r in enumerate(body):
        name = col(r, "Kernel Name").split("(")[0].split("::")[-1]
        du = float(col(r, "gpu__time_duration.sum").replace(",", ""))
        uu = units[hdr.index("gpu__time_duration.sum")]
        du = du / 1e3 if uu in ("ns", "nsecond") else du * 1e3 if uu in ("ms", "msecond") else du
        rd = to_mb(col(r, "dram__bytes_read.sum"), units[hdr.index("dram__bytes_read.sum")])
        wr = to_mb(col(r, "dram__bytes_write.sum"), units[hdr.index("dram__bytes_write.sum")])
        tb += rd + wr
        tt += du
        out.append(f"| {i} | `{name}` ({col(r, 'launch__grid_size')}) | {du:.1f} | {rd:.0f} | {wr:.0f} | "
                   f"{col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')[:5]} | "
                   f"{col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')[:5]} | "
                   f"{col(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')[:5]} | "
                   f"{col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[:5]} |")
    n = max(1, len(body))
    out.append(f"\naverage DRAM traffic per launch: {tb / n:.1f} MB; average duration {tt / n:.1f} us\n")
    return "\n".join(out)


def main():
    bench, den, dec, gemm, conv = sys.argv[1:6]
    light = sys.argv[6] if len(sys.argv) > 6 else None
    b = json.loads([l for l in read(bench).splitlines() if l.startswith("{")][-1])
    r = b["roofline"]
    rows = [dict(r, **{"class": "gemm_tc"})] + r["secondary"]
    o = ["# r02 — per-kernel roofline evidence (final build of round 2)\n",
         "Two sources, same box class (1 x B200):\n",
         "1. **In-step, live** (`bench.py`, CUDA events around every launch of one more AR step after the timed region): "
         "algorithmic FLOPs / bytes per class divided by the class's summed launch time. These are the numbers the roofline "
         f"fractions use (SM clock {b['clocks']['sm_mhz']:.0f} MHz median under {', '.join(b['clocks']['reasons']) or 'no throttle reason'}).",
         "2. **ncu `--clock-control none`** of the same kernels at the production shapes (`tools/prof_all.py`: one 375M denoiser call "
         "B=20 T_out=4, one 80-frame decode, metrics at 20 / 50 members, Heun step, attention B=13 H=16): DRAM bytes read+written "
         "per launch, pipe utilisation, stall reasons (raw pages: `r02_ncu_*_raw.csv.gz`). ncu serialises launches with a cold L2, "
         "so its durations are upper bounds of the in-step ones. The `--set full` captures (sections 2-4) predate the staged GEMM "
         "epilogue; section 5 is a light re-capture of the GEMM launches with it.\n",
         f"Step: {b['ms_per_step']:.1f} ms per AR step (20 members x 4 leads: {b['value']:.1f} member-steps/s device-resident, "
         f"{b['e2e']['value']:.1f} end to end with H2D noise + D2H fields), whole-step algorithmic rate {r['step_algorithmic_tflops']} TFLOP/s.\n",
         "## 1. In-step class table (bench.py)\n",
         "| class | kernel | launches | ms | share | bound | achieved | peak | frac | algorithmic MB / launch |",
         "|---|---|---:|---:|---:|---|---:|---:|---:|---:|"]
    for x in rows:
        ab = x.get("algorithmic_bytes")
        o.append(f"| {x['class']} | {x['kernel'].split(' (')[0]} | {x['launches']} | {x['ms']:.2f} | {100 * x['share_of_step']:.1f}% | {x['bound']} | "
                 f"{x['achieved']} {x['unit']} | {x['peak']} | **{x['frac']:.3f}** | {'' if ab is None else round(ab / x['launches'] / 1e6, 1)} |")
    tot = sum(x["ms"] for x in rows)
    o.append(f"\nSum of classes {tot:.1f} ms of the {b['ms_per_step']:.1f} ms step; the rest is launch boundaries "
             f"({b['gpu_launches']} launches in {b['steps']} steps).\n")
    if b.get("metrics"):
        o += ["Metrics kernel (`metrics_sorted_kernel<M>`, once per evaluated AR step, decoded fields [M, 84, 4, 120, 240]):\n",
              "| members | kernel ms | algorithmic MB | GB/s | frac of the HBM copy peak |", "|---:|---:|---:|---:|---:|"]
        for m in b["metrics"]:
            o.append(f"| {m['members']} | {m['kernel_ms']} | {m['algorithmic_bytes'] / 1e6:.0f} | {m['achieved']} | {m['frac']} |")
        o.append("")
    rep = os.path.join(ROOT, "tools", "ncu_kernel_report.py")
    o.append("## 2. ncu, non-GEMM kernels\n")
    o.append(subprocess.run([sys.executable, rep, plain(den), plain(dec)], capture_output=True, text=True).stdout)
    o.append("\n## 3. ncu, tcgen05 GEMM (first 40 launches of a denoiser call) and implicit-GEMM sphere conv (first 24 GEMM launches of a decode)\n")
    o.append(subprocess.run([sys.executable, rep, plain(gemm), plain(conv)], capture_output=True, text=True).stdout)
    o.append("\n## 4. Per-launch GEMM / conv tables (direct-store epilogue, before the staged write-back)\n")
    o.append(launch_table(gemm, "Denoiser call (375M, B=20, T_out=4): first 40 GEMM launches, in call order"))
    o.append(launch_table(conv, "DC-AE decode (80 frames): first 24 GEMM-kernel launches (3x3 sphere convs in implicit-GEMM mode and 1x1 GEMMs), in order"))
    if light:
        o.append("\n## 5. Per-launch GEMM table with the staged bf16 epilogue (final build; light metric set)\n")
        o.append(launch_table(light, "Denoiser call (375M, B=20, T_out=4): first 40 GEMM launches, in call order"))
    print("\n".join(o))


if __name__ == "__main__":
    main()
