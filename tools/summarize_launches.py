"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (markdown)."""
import collections
import csv
import re
import sys


def main(path, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        name = name.split("::")[-1] if "lc::" in name else name[:70]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {title}\n")
    print(f"`ncu --metrics gpu__time_duration.sum --clock-control none` over one AR step of `bench.py` (ens=20, 375M, 20 "
          f"denoise calls, T_out=4, decode of 80 frames): {sum(cnt.values())} launches, {total:.1f} ms summed kernel time "
          f"(cold-cache, serialised: compare shares, not absolutes).\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        if v / total < 0.0005:
            continue
        print(f"| `{k}` | {cnt[k]} | {v:.2f} | {100 * v / total:.1f}% | {1e3 * v / cnt[k]:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "launch list")
