"""Turn .ncu-rep files / ncu launch lists into the small CSV summaries committed under profiles/.
  python scripts/ncu_summary.py rep  <file.ncu-rep> <out.csv>     # selected raw metrics, one block per captured launch
  python scripts/ncu_summary.py list <launches.csv> <out.csv>     # per-kernel totals and shares of a launch list"""
import collections
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "lts__t_bytes.sum",
        "lts__throughput.avg.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__", "sm__throughput.avg.pct",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "smsp__inst_executed.sum", "sm__warps_active.avg.pct",
        "smsp__issue_active.avg.pct", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__inst_executed_pipe_xu",
        "l1tex__t_bytes.sum", "smsp__average_warp", "sm__cycles_active.avg")


def rep(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            f.write("## " + name[:110] + "\n")
            for h, u, v in sorted(zip(hdr, units, vals)):
                if any(h.startswith(k) for k in KEEP) and "Triage" not in h and ".max" not in h and ".min" not in h:
                    f.write(f"{h},{u},{v}\n")


def lst(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v for _, v in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_ms,share\n")
        for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k},{c},{v / 1e6:.3f},{v / tot:.4f}\n")


if __name__ == "__main__":
    {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2], sys.argv[3])
