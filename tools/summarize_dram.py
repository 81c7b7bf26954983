"""DRAM bytes per kernel from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list (the second half of profiles/r02_launches_dram.txt). Usage: python tools/summarize_dram.py launches.csv"""
import collections
import csv
import re
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    m = r.get("Metric Name", "")
    if not m.startswith("dram__bytes"):
        continue
    k = re.sub(r"\(.*", "", r["Kernel Name"])
    a = agg.setdefault(k, {"ids": set(), "read": 0.0, "write": 0.0})
    a["ids"].add(r["ID"])
    a["read" if "read" in m else "write"] += float(r["Metric Value"].replace(",", "")) * SCALE.get(r.get("Metric Unit", "byte"), 1.0)
rd = sum(a["read"] for a in agg.values())
wr = sum(a["write"] for a in agg.values())
n = sum(len(a["ids"]) for a in agg.values())
print(f"DRAM traffic of the step (dram__bytes_read.sum + dram__bytes_write.sum over all {n} launches): "
      f"read {rd / 1e9:.2f} GB + write {wr / 1e9:.2f} GB = {(rd + wr) / 1e9:.2f} GB")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1]["read"] + kv[1]["write"])):
    if a["read"] + a["write"] < 5e7:
        continue
    print(f"  {len(a['ids']):5d} x  read {a['read'] / 1e9:7.3f} GB  write {a['write'] / 1e9:7.3f} GB  {k}")
