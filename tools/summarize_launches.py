"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (profiles/*.md)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None, title=""):
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        name = name.split("::")[-1]
        unit = r.get("Metric Unit", "ns")
        v = float(r["Metric Value"].replace(",", ""))
        if unit in ("us", "usecond"):
            v *= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e6
        rows.append((name, v))
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        agg[n][0] += 1
        agg[n][1] += v
    total = sum(v for _, v in rows) or 1.0
    lines = ["# %s" % (title or path), "",
             "ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`): per-launch times are cold-cache and",
             "serialised — compare SHARES, not absolutes.", "", "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (n, c, t / 1e3, t / c / 1e3, 100 * t / total))
    lines.append("| **total** | %d | %.1f | | 100%% |" % (len(rows), total / 1e3))
    txt = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "")
