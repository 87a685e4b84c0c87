#!/usr/bin/env python
"""Average DRAM bytes per zgemm launch from an ncu metrics CSV of the default bench command
(ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --kernel-name-base demangled -k regex:zgemm --csv ...)
-> profiles/r01_zgemm_traffic.json, which bench.py reports as roofline.traffic."""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ni, vi, ui, ii = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = {}
for r in rows[1:]:
    if r[ni].startswith("dram__bytes"):
        per[r[ii]] = per.get(r[ii], 0.0) + float(r[vi].replace(",", "")) * mult.get(r[ui], 1)
vals = list(per.values())
out = {"workload": sys.argv[2], "solves_per_step_per_gpu": int(sys.argv[3]), "launches_captured": len(vals),
       "dram_bytes_per_launch": sum(vals) / len(vals), "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the zgemm launches of the default bench command"}
json.dump(out, open(sys.argv[4], "w"), indent=1)
print(out)
