"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (profiles/*.md)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None, title="", step=None):
    """step = k: keep only the launches of the k-th train step (1-based; a step ends with its adam_amsgrad_kernel)."""
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
    if step is not None:
        ends = [i for i, (n, _) in enumerate(rows) if n.startswith("adam_amsgrad_kernel")]
        lo = ends[step - 2] + 1 if step >= 2 else 0
        rows = rows[lo:ends[step - 1] + 1]
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
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "",
         int(sys.argv[4]) if len(sys.argv) > 4 else None)
