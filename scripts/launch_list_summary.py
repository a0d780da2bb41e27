"""Per-kernel summary of an ncu launch list (csv from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv --log-file X.csv python bench.py ...`): launches, total / mean time, share of the listed time, DRAM MB per launch.
usage: python scripts/launch_list_summary.py gpurun_out/launches.csv "header comment" > profiles/rNN_launch_list_summary.csv"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
t, rd, wr = collections.defaultdict(list), collections.defaultdict(float), collections.defaultdict(float)
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
for r in rows[1:]:
    k = r[ki].split("(")[0].replace("void ", "")
    v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    if r[mi] == "gpu__time_duration.sum":
        t[k].append(v)
    elif r[mi] == "dram__bytes_read.sum":
        rd[k] += v
    elif r[mi] == "dram__bytes_write.sum":
        wr[k] += v
tot = sum(sum(v) for v in t.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("kernel,launches,total_us,mean_us,share,dram_read_MB_per_launch,dram_write_MB_per_launch")
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    n = len(v)
    print("%s,%d,%.1f,%.1f,%.3f,%.1f,%.1f" % (k, n, sum(v) / 1e3, sum(v) / n / 1e3, sum(v) / tot, rd[k] / n / 1e6, wr[k] / n / 1e6))
