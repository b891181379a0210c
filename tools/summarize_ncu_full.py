"""Summarise an `ncu --set full` report (.ncu-rep) into a small markdown table for profiles/."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def main(rep, out=None, title=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units = rd[0], rd[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# %s" % (title or rep), "", "`ncu --set full --clock-control none --import-source on` (one GPU, cold caches per replay).", ""]
    seen = {}
    for row in rd[2:]:
        kn = row[idx["Kernel Name"]].strip()
        seen[kn] = seen.get(kn, 0) + 1
        if seen[kn] > 2:
            continue
        lines.append("## %s" % kn)
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---:|---|")
        for w in WANT:
            if w in idx:
                lines.append("| `%s` | %s | %s |" % (w, row[idx[w]], units[idx[w]]))
        lines.append("")
    txt = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(txt)
    print(txt[:3000])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "")
